"""Cycle accounting of K2's pipeline roles (library built with CNB_NVCC_EXTRA=-DCNB_TRACE)."""
import os, sys, ctypes, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import codenerf_b200 as cn
from codenerf_b200 import synthetic as syn, ops, _lib
from tests import gpu_util as U
model, flat = U.make_model("bf16")
L = _lib.load()
N, n_seg, R = 64, 32, 2048
c2ws = np.stack([syn.look_at_pose(700 + g, 1.3) for g in range(n_seg)])
zs = np.stack([np.linspace(0.8, 1.8, N).astype(np.float32) for g in range(n_seg)])
bundle = cn.RayBundle(z_vals=torch.from_numpy(zs).cuda(), rays_per_segment=R, c2w=torch.from_numpy(c2ws).cuda(),
                      pix_begin=torch.zeros(n_seg, dtype=torch.int32).cuda(), focal=torch.tensor([131.25], dtype=torch.float64), H=128, W=128)
sc = torch.from_numpy(syn.make_codes(1, n_seg)).cuda(); tc = torch.from_numpy(syn.make_codes(2, n_seg)).cuda()
params = model.param_list(); packed = model._packed.get(model._cfg, params)
rb = bundle.args(sc, tc)
d_rgb = torch.randn(n_seg * R, 3, device="cuda") * 1e-4
buf = (ctypes.c_ulonglong * 32)()
names = ["mma.wait_a_ready", "mma.wait_w_full", "mma.total", "auxX.wait_ready", "auxX.total", "X.wait_buf_free", "X.wait_acc_fwd",
         "X.epilogue_fwd(+buf)", "X.composite+step0", "X.wait_acc_bwd", "X.epilogue_bwd(+buf)", "X.encode", "X.total"]
for want in (True, False):
    for _ in range(2): ops.render_backward(model._cfg, params, packed, rb, 0, d_rgb, None, want)
    torch.cuda.synchronize()
    L.cnb_debug_trace_bwd(None, 1)
    iters = 3
    for _ in range(iters): ops.render_backward(model._cfg, params, packed, rb, 0, d_rgb, None, want)
    torch.cuda.synchronize()
    L.cnb_debug_trace_bwd(buf, 1)
    tiles = n_seg * R * N / 128
    print(f"param grads={want}: cycles per CTA per step, and per tile-pair-op (17 GEMM ops per tile)")
    for i, n in enumerate(names):
        per_cta = buf[i] / iters / 148
        print(f"  {n:24s} {per_cta:12.0f}   per tile-pair-op {per_cta / (tiles / 148 / 2) / 17:8.0f}")
