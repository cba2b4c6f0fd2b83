import sys, time, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import codenerf_b200 as cn
from codenerf_b200 import synthetic as syn
from tests import gpu_util as U
model, flat = U.make_model("bf16")
for N, n_seg in ((64, 32), (64, 64), (96, 32)):
    R = 2048
    c2ws = np.stack([syn.look_at_pose(700 + g, 1.3) for g in range(n_seg)])
    zs = np.stack([np.linspace(0.8, 1.8, N).astype(np.float32) for g in range(n_seg)])
    pix = np.zeros(n_seg, np.int32)
    bundle = cn.RayBundle(z_vals=torch.from_numpy(zs).cuda(), rays_per_segment=R, c2w=torch.from_numpy(c2ws).cuda(),
                          pix_begin=torch.from_numpy(pix).cuda(), focal=torch.tensor([131.25], dtype=torch.float64), H=128, W=128)
    sc = torch.from_numpy(syn.make_codes(1, n_seg)).cuda(); tc = torch.from_numpy(syn.make_codes(2, n_seg)).cuda()
    with torch.no_grad():
        for _ in range(3): cn.render(model, bundle, sc, tc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters): cn.render(model, bundle, sc, tc)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    rays = n_seg * R
    flops = rays * N * 899328
    print(f"N={N} n_seg={n_seg}: {ms:.3f} ms  {rays/ms/1e3:.3f} Mrays/s  {flops/ms/1e9:.1f} TFLOP/s ({flops/ms/1e9/1651.9*100:.1f}% of measured bf16 peak)")
