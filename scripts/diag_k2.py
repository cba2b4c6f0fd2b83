"""K2 pipeline diagnosis on the instrumented build (CNB_LIB=trace; `python -m codenerf_b200.build --trace`).

For each mode (default / CTA pairs) x (training / latent fit): kernel time, cycle accounting of the pipeline
roles and the per-op timeline of CTA 0 (operands ready -> MMAs issued -> accumulator visible -> epilogue done).
"""
import os, sys, ctypes
os.environ.setdefault("CNB_LIB", "trace")
import torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import codenerf_b200 as cn
from codenerf_b200 import synthetic as syn, ops, _lib
from tests import gpu_util as U

model, flat = U.make_model("bf16")
L = _lib.load()
N, n_seg, R = 64, 16, 2048
c2ws = np.stack([syn.look_at_pose(700 + g, 1.3) for g in range(n_seg)])
zs = np.stack([np.linspace(0.8, 1.8, N).astype(np.float32) for g in range(n_seg)])
bundle = cn.RayBundle(z_vals=torch.from_numpy(zs).cuda(), rays_per_segment=R, c2w=torch.from_numpy(c2ws).cuda(),
                      pix_begin=torch.zeros(n_seg, dtype=torch.int32).cuda(), focal=torch.tensor([131.25], dtype=torch.float64), H=128, W=128)
sc = torch.from_numpy(syn.make_codes(1, n_seg)).cuda(); tc = torch.from_numpy(syn.make_codes(2, n_seg)).cuda()
params = model.param_list(); packed = model._packed.get(model._cfg, params)
rb = bundle.args(sc, tc)
tgt = torch.rand(n_seg * R, 3, device="cuda")
nl, n_ops = 8, 15
names = ["mma.wait_a_ready", "mma.wait_w_full", "mma.total", "auxX.wait_ready", "auxX.total", "X.wait_buf_free", "X.wait_acc_fwd",
         "X.epilogue_fwd(+buf)", "X.composite+step0", "X.wait_acc_bwd", "X.epilogue_bwd(+buf)", "X.encode", "X.total", "X.mid: composite part", "X.mid: head MMA wait"]
fbuf = (ctypes.c_float * 512)()
tbuf = (ctypes.c_ulonglong * 32)()
ev = (ctypes.c_ulonglong * (4 * 16384))(); cnt = (ctypes.c_uint * 4)()
tiles = n_seg * R * N / 128


dP = torch.zeros(sum(p.numel() for p in params), device="cuda")


def step(want):
    return ops.render_train_step(model._cfg, params, packed, rb, 0, tgt, 1.0, dP if want else None, want_outputs=False)


def mean(x):
    return float(np.mean(x)) if len(x) else float("nan")


modes = [s for s in os.environ.get("DIAG_MODES", "0,1").split(",")]
for pairs in modes:
    os.environ["CNB_BWD_PAIRS"] = pairs
    for want in (True, False):
        for _ in range(2): step(want)
        torch.cuda.synchronize(); L.cnb_profile_enable(1)
        L.cnb_debug_trace_bwd(None, 1)
        iters = 3
        for _ in range(iters): step(want)
        torch.cuda.synchronize()
        L.cnb_debug_trace_bwd(tbuf, 1)
        kt = {}
        for kid, name in ((1, "bwd"), (2, "wgrad")):
            n = L.cnb_profile_read(kid, fbuf, 512)
            if n > 0: kt[name] = round(sum(fbuf[i] for i in range(n)) / iters, 3)
        L.cnb_profile_enable(0)
        print(f"\n=== pairs={pairs} param_grads={want}: kernel ms per step {kt}; timeouts {L.cnb_debug_pipeline_timeouts()}")
        print("  cycles per CTA per step | per tile")
        for i, nm in enumerate(names):
            per_cta = tbuf[i] / iters / 148
            print(f"  {nm:24s} {per_cta:12.0f} {per_cta / (tiles / 148):10.0f}")
        L.cnb_debug_events_bwd(None, None, 1)
        step(want)
        torch.cuda.synchronize()
        L.cnb_debug_events_bwd(ev, cnt, 1)
        E = {}
        for slot in range(3):
            n = min(cnt[slot], 16384)
            for i in range(n):
                v = ev[slot * 16384 + i]
                t, code = v >> 16, v & 0xffff
                E.setdefault((code >> 12, (code >> 8) & 0xf, code & 0xff), []).append(t)
        print("  op | g | ready->issued | ready->acc | epilogue | epi end->next ready | period (ready->next tile same op)")
        for g in (0, 1):
            for op in range(n_ops):
                ready = E.get((1, g, op), []); issued = E.get((2, g, op), []); acc = E.get((3, g, op), []); done = E.get((4, g, op), [])
                if op == nl - 1: done = E.get((5, g, nl), done)
                nxt = E.get((1, g, op + 1), [])
                n = min(len(ready), len(issued), len(acc), len(done))
                if n == 0: continue
                r, i_, a, d = (np.array(x[:n], dtype=np.float64) for x in (ready, issued, acc, done))
                line = f"  {op:2d} | {g} | {mean(i_ - r):7.0f} | {mean(a - r):7.0f} | {mean(d - a):7.0f}"
                m = min(n, len(nxt))
                line += f" | {mean(np.array(nxt[:m], dtype=np.float64) - d[:m]):7.0f}" if m else " |     nan"
                if n > 2: line += f" | {mean(np.diff(r)):8.0f}"
                print(line)
        # tile period of group 0: PE store to PE store
        st = np.array(E.get((6, 0, 0), []), dtype=np.float64)
        if len(st) > 2: print(f"  group X tile period (cycles): {mean(np.diff(st)):.0f}  over {len(st)} tiles")
        sys.stdout.flush()
