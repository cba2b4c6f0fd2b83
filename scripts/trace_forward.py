"""Cycle accounting of K1's pipeline roles (library built with CNB_NVCC_EXTRA=-DCNB_TRACE)."""
import os, sys, ctypes, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import codenerf_b200 as cn
from codenerf_b200 import synthetic as syn, _lib
from tests import gpu_util as U
model, flat = U.make_model("bf16")
L = _lib.load()
N, n_seg, R = 64, 32, 2048
c2ws = np.stack([syn.look_at_pose(700 + g, 1.3) for g in range(n_seg)])
zs = np.stack([np.linspace(0.8, 1.8, N).astype(np.float32) for g in range(n_seg)])
bundle = cn.RayBundle(z_vals=torch.from_numpy(zs).cuda(), rays_per_segment=R, c2w=torch.from_numpy(c2ws).cuda(),
                      pix_begin=torch.zeros(n_seg, dtype=torch.int32).cuda(), focal=torch.tensor([131.25], dtype=torch.float64), H=128, W=128)
sc = torch.from_numpy(syn.make_codes(1, n_seg)).cuda(); tc = torch.from_numpy(syn.make_codes(2, n_seg)).cuda()
buf = (ctypes.c_ulonglong * 32)()
names_ss = ["prod.wait_empty", "prod.total", "mma.wait_a_ready", "mma.wait_w_full", "mma.total",
            "X.wait_acc", "X.epilogue", "X.encode", "X.total", "Y.wait_acc", "Y.epilogue", "Y.encode", "Y.total"]
names_ts = ["-", "-", "mma.wait_operand", "mma.wait_w_full", "mma.total", "epi.wait_d_full0", "epi.wait_d_full1",
            "epi.finalize", "epi.total", "io.wait_pe_free", "io.wait_samples", "io.total", "-", "mma.wait_d_free"]
for form in ("ts", "ss"):
    if len(sys.argv) > 1 and form not in sys.argv[1:]: continue
    os.environ["CNB_FWD_KERNEL"] = form
    names = names_ts if form == "ts" else names_ss
    mc = form
    with torch.no_grad():
        for _ in range(3): cn.render(model, bundle, sc, tc)
        torch.cuda.synchronize()
        L.cnb_debug_trace(None, 1)
        iters = 5
        for _ in range(iters): cn.render(model, bundle, sc, tc)
        torch.cuda.synchronize()
    L.cnb_debug_trace(buf, 1)
    tiles = n_seg * R * N / 128
    print(f"kernel form {mc}: cycles per CTA per launch (avg over 148 CTAs), and per tile-layer (9 layers)")
    for i, n in enumerate(names):
        per_cta = buf[i] / iters / 148
        print(f"  {n:18s} {per_cta:12.0f}   per tile-pair-layer {per_cta / (tiles / 148 / 2) / 9:8.0f}")
