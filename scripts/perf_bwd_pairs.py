"""K2 on CTA pairs (CNB_BWD_PAIRS=1): timings and agreement with the default kernel."""
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
buf = (ctypes.c_float * 512)()
ref = {}
for pairs in (0, 1):
    os.environ["CNB_BWD_PAIRS"] = str(pairs)
    line = f"pairs={pairs}:"
    for want in (True, False):
        for _ in range(2): res = ops.render_backward(model._cfg, params, packed, rb, 0, d_rgb, None, want)
        torch.cuda.synchronize(); L.cnb_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): res = ops.render_backward(model._cfg, params, packed, rb, 0, d_rgb, None, want)
        e1.record(); torch.cuda.synchronize()
        kt = {}
        for kid, name in ((1, "bwd"), (2, "wgrad")):
            n = L.cnb_profile_read(kid, buf, 512)
            if n > 0: kt[name] = round(sum(buf[i] for i in range(n)) / 5, 3)
        L.cnb_profile_enable(0)
        ms = e0.elapsed_time(e1) / 5
        line += f" | grads={want}: {ms:.3f} ms {n_seg*R/ms/1e3:.2f} Mrays/s {kt}"
        tens = [t for t in res if isinstance(t, torch.Tensor)]
        key = f"bwd{want}"
        if pairs == 0: ref[key] = [t.clone() for t in tens]
        else:
            err = max(((a - b).abs().max() / (b.abs().max() + 1e-30)).item() for a, b in zip(tens, ref[key]) if a.numel())
            line += f" maxrel vs default {err:.2e}"
    print(line, "timeouts", L.cnb_debug_pipeline_timeouts(), flush=True)
