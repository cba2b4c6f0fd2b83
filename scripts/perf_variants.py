"""Forward kernel variants: timing and agreement with the default kernel."""
import os, sys, torch, numpy as np
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
variants = [("default", {}), ("epilogue8", {"CNB_EPI_WARPS": "8"}), ("pairs", {"CNB_CTA_PAIRS": "1"}),
            ("pairs+epilogue8", {"CNB_CTA_PAIRS": "1", "CNB_EPI_WARPS": "8"})]
ref = None
for name, env in variants:
    for k in ("CNB_EPI_WARPS", "CNB_CTA_PAIRS"): os.environ.pop(k, None)
    os.environ.update(env)
    with torch.no_grad():
        for _ in range(3): out = cn.render(model, bundle, sc, tc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): out = cn.render(model, bundle, sc, tc)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    rays = n_seg * R
    if ref is None: ref = [o.clone() for o in out]; err = 0.0
    else: err = max((a - b).abs().max().item() for a, b in zip(out, ref))
    print(f"{name:16s}: {ms:.3f} ms {rays/ms/1e3:.2f} Mrays/s ({rays*N*899328/ms/1e9/1651.9*100:.1f}% peak)  max|diff vs default| {err:.2e}  timeouts {L.cnb_debug_pipeline_timeouts()}", flush=True)
