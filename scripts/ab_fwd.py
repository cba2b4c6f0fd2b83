"""Same-session A/B of forward-kernel options: python scripts/ab_fwd.py "" CNB_EPI_WARPS=8 ..."""
import os, sys, subprocess, json
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, json, torch, numpy as np
sys.path.insert(0, %r)
import bench, codenerf_b200 as cn
from codenerf_b200 import synthetic as syn, ops, _lib
from tests import gpu_util as U
model, flat = U.make_model("bf16")
c2w, pix, z, tgt, sc, tc = bench.synthetic_batch(32, 0)
T = lambda a: torch.from_numpy(a).cuda()
b = cn.RayBundle(z_vals=T(z), rays_per_segment=2048, c2w=T(c2w), pix_begin=T(pix), focal=torch.tensor([131.25], dtype=torch.float64), H=128, W=128)
params = model.param_list(); packed = model._packed.get(model._cfg, params)
rb = b.args(T(sc), T(tc))
for _ in range(5): ops.render_forward(model._cfg, params, packed, rb, 0)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): ops.render_forward(model._cfg, params, packed, rb, 0)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 30)
print(json.dumps({"ms": best, "Mrays": 65536 / best / 1e3, "timeouts": _lib.load().cnb_debug_pipeline_timeouts()}))
''' % root
for rnd in range(2):
    for v in (sys.argv[1:] or [""]):
        env = dict(os.environ)
        for part in v.split(","):
            if "=" in part: k, x = part.split("=", 1); env[k] = x
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=root)
        print(f"{v or 'default':40s}", out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)
