"""Timeline of CTA 0 of K2 (CNB_TRACE build): per GEMM op, when operands are ready, MMAs issued, accumulator visible,
epilogue done.  Prints mean durations per op and the overlap structure."""
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
want = len(sys.argv) > 1 and sys.argv[1] == "train"
for _ in range(2): ops.render_backward(model._cfg, params, packed, rb, 0, d_rgb, None, want)
torch.cuda.synchronize()
ev = (ctypes.c_ulonglong * (4 * 16384))(); cnt = (ctypes.c_uint * 4)()
L.cnb_debug_events_bwd(None, None, 1)
ops.render_backward(model._cfg, params, packed, rb, 0, d_rgb, None, want)
torch.cuda.synchronize()
L.cnb_debug_events_bwd(ev, cnt, 1)
E = {}
for slot in range(3):
    n = min(cnt[slot], 16384)
    for i in range(n):
        v = ev[slot * 16384 + i]
        t, code = v >> 16, v & 0xffff
        kind, g, op = code >> 12, (code >> 8) & 0xf, code & 0xff
        E.setdefault((kind, g, op), []).append(t)
nl = 9; n_ops = 17
print("events per slot:", list(cnt)[:3], "(first launch of the step only: the counters stop at 16384)")
def mean(x): return float(np.mean(x)) if len(x) else float('nan')
print("op | kind | g | ready->issued | ready->acc visible | epilogue | epilogue end -> next ready (sync) ")
for g in (0, 1):
    for op in range(n_ops):
        ready = E.get((1, g, op), []); issued = E.get((2, g, op), []); acc = E.get((3, g, op), []); done = E.get((4, g, op), [])
        if op == nl - 1: done = E.get((5, g, nl), done)          # last forward layer: the epilogue end is after compositing + step 0
        nxt = E.get((1, g, op + 1), [])
        n = min(len(ready), len(issued), len(acc), len(done))
        if n == 0: continue
        r, i_, a, d = (np.array(x[:n], dtype=np.float64) for x in (ready, issued, acc, done))
        line = f"{op:2d} | {'fwd' if op < nl else 'bwd'} | {g} | {mean(i_ - r):7.0f} | {mean(a - r):7.0f} | {mean(d - a):7.0f}"
        m = min(n, len(nxt))
        if m: line += f" | {mean(np.array(nxt[:m], dtype=np.float64) - d[:m]):7.0f}"
        print(line)
# pipe sharing: how long after X's ready does its issue start relative to Y's issue end
