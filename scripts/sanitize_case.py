"""A 2-tile problem through K1 (fused forward), K2 (recompute + input gradients, with and without the stash), K3 (weight
gradients) and the head kernel, for compute-sanitizer (one tool per run):

    compute-sanitizer --tool racecheck|synccheck|memcheck python scripts/sanitize_case.py

Small enough that the instrumented kernels finish in seconds; checks the results against the oracle as well."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import codenerf_b200 as cn
from codenerf_b200 import synthetic as syn, ops, _lib
from oracle import oracle as orc
from tests import gpu_util as U

for N, rays in ((64, 4), (96, 4)):
    model, flat = U.make_model("bf16")
    H = W = 16
    focal = 131.25 * W / 128
    c2w = syn.look_at_pose(3, 1.3)
    z = orc.z_vals(0.8, 1.8, N, orc.torch_rand(5, N))
    sc, tc = syn.make_codes(11, 1), syn.make_codes(12, 1)
    tgt = syn.make_targets(13, rays)
    ref = orc.render(flat, H, W, focal, c2w, z, sc, tc, True, ray_begin=7, ray_count=rays)
    bundle = cn.RayBundle(z_vals=torch.from_numpy(z[None]).cuda(), rays_per_segment=rays, c2w=torch.from_numpy(c2w[None]).cuda(),
                          pix_begin=torch.tensor([7], dtype=torch.int32).cuda(), focal=torch.tensor([focal], dtype=torch.float64), H=H, W=W)
    params = model.param_list(); packed = model._packed.get(model._cfg, params)
    rb = bundle.args(torch.from_numpy(sc).cuda(), torch.from_numpy(tc).cuda())
    rgb, depth, acc = ops.render_forward(model._cfg, params, packed, rb, 0)                      # K1
    dP = torch.zeros(flat.size, device="cuda")
    out = ops.render_train_step(model._cfg, params, packed, rb, 0, torch.from_numpy(tgt).cuda(), 1.0, dP)   # K2 + K3 + head
    out2 = ops.render_train_step(model._cfg, params, packed, rb, 0, torch.from_numpy(tgt).cuda(), 1.0, None)  # K2, no stash
    torch.cuda.synchronize()
    assert np.abs(rgb.cpu().numpy() - ref["rgb"]).max() < 1e-2 and np.abs(out[0].cpu().numpy() - ref["rgb"]).max() < 1e-2
    assert torch.isfinite(dP).all() and float(dP.abs().max()) > 0
    # (the latent fit's code gradients come from the auxiliary warps' packed-bf16 pre-sums, the training step's from K3's fp32 sums)
    assert float((out[4] - out2[4]).abs().max()) <= 2e-2 * float(out2[4].abs().max())
    print(f"N={N}: K1/K2/K3 ran, max|rgb - oracle| = {np.abs(rgb.cpu().numpy() - ref['rgb']).max():.2e}, timeouts {_lib.load().cnb_debug_pipeline_timeouts()}")
