#!/usr/bin/env python
"""bench.py -- throughput of the CodeNeRF render path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): srncar.json network, one training step over a batch of
`--objects` objects x 2048 rays x 64 samples per GPU: fused forward (ray gen + sampling + PE +
MLP + compositing) + L2 loss + backward (weight, bias and latent-code gradients), synthetic
SRN-shaped data (128x128 views, random-init weights and codes).  For N > 1 each rank runs the
same per-GPU batch on its own objects (weak scaling) and the flat MLP gradient is all-reduced
with NCCL every step (configs[3]).  One JSON line is printed by rank 0.

`value` times the step with inputs resident in HBM; `e2e` times the same step through the
public API with HOST inputs (pinned poses / z_vals / target pixels copied in, per-segment loss
copied out, every step).  `--impl reference` times the CPU restatement of the reference
(oracle/, C + OpenMP on all host cores) on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rays/sec (64 samples/ray) fwd+bwd training step"
UNIT = "rays/s"
N_SAMPLES = 64
RAYS_PER_OBJECT = 2048                 # reference train.py:17 batch_size
FLOP_FWD = 899_328                     # per sample, SURVEY.md 8(d)
FLOP_TRAIN = 2_651_904
FLOP_DGRAD = 853_248


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained"),
                    hbm_gbs=d["hbm_gbs"], source="measured")
    return dict(bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_batch(n_obj, rank):
    from codenerf_b200 import synthetic as syn
    base = 100000 * rank
    c2w = np.stack([syn.look_at_pose(base + g, syn.SRN_CARS["radius"]) for g in range(n_obj)])
    pix = np.array([(2048 * ((base + g) % 8)) for g in range(n_obj)], np.int32)       # one of the view's 8 chunks
    near, far = syn.SRN_CARS["near"], syn.SRN_CARS["far"]
    dist = (far - near) / (2 * N_SAMPLES)
    z0 = np.linspace(near + dist, far - dist, N_SAMPLES, dtype=np.float32)
    jit = syn.uniform(base + 5, n_obj * N_SAMPLES).reshape(n_obj, N_SAMPLES).astype(np.float32)
    z = (z0[None] + jit * np.float32(far - near) / np.float32(2 * N_SAMPLES)).astype(np.float32)
    tgt = syn.make_targets(base + 9, n_obj * RAYS_PER_OBJECT)
    sc, tc = syn.make_codes(base + 11, n_obj), syn.make_codes(base + 12, n_obj)
    return c2w, pix, z, tgt, sc, tc


# ---------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's CPU implementation of the path (oracle port, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from codenerf_b200 import synthetic as syn
    from oracle import oracle as orc
    flat, _ = syn.make_params(0)
    n_rays = args.ref_rays
    c2w, pix, z, tgt, sc, tc = synthetic_batch(1, 0)

    def step():
        fwd = orc.render(flat, 128, 128, syn.SRN_FOCAL, c2w[0], z[0], sc[:1], tc[:1], True, ray_begin=0, ray_count=n_rays)
        d_rgb = (2.0 * (fwd["rgb"] - tgt[:n_rays]) / (3.0 * n_rays)).astype(np.float32)
        orc.render_backward(flat, fwd, z[0], sc[:1], tc[:1], d_rgb, None, True)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = n_rays / dt
    sample = f"{n_rays} rays x {N_SAMPLES} samples of one 2048-ray chunk per step, fwd+bwd incl. weight gradients"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"srncar.json training step: {args.objects} objects x {RAYS_PER_OBJECT} rays x {N_SAMPLES} samples "
                                   "per GPU, fused forward + L2 loss + backward (weights, biases, codes)",
                       "net": "W=256, 3 shape + 1 texture blocks, latent 256", "view": "128x128", "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": orc.num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(n_rays=512):
    """Bounded CPU sample timed beside the GPU number (rank 0, N=1 only)."""
    from codenerf_b200 import synthetic as syn
    from oracle import oracle as orc
    flat, _ = syn.make_params(0)
    c2w, pix, z, tgt, sc, tc = synthetic_batch(1, 0)
    best = None
    for it in range(2):
        t0 = time.perf_counter()
        fwd = orc.render(flat, 128, 128, syn.SRN_FOCAL, c2w[0], z[0], sc[:1], tc[:1], True, ray_begin=0, ray_count=n_rays)
        d_rgb = (2.0 * (fwd["rgb"] - tgt[:n_rays]) / (3.0 * n_rays)).astype(np.float32)
        orc.render_backward(flat, fwd, z[0], sc[:1], tc[:1], d_rgb, None, True)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": n_rays / best, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
            "sample": f"{n_rays} rays x {N_SAMPLES} samples fwd+bwd, best of 2, oracle/ C port with OpenMP"}


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import codenerf_b200 as cn
    from codenerf_b200 import _lib, ops
    from codenerf_b200 import synthetic as syn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.load()
    _lib.check(L.cnb_check_device())

    n_obj = args.objects
    cfg_net = dict(syn.SRN_NET)
    flat, views = syn.make_params(0, cfg_net)
    model = cn.CodeNeRF(**cfg_net, precision=args.precision)
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in views.items()})
    model = model.to(dev)
    params = model.param_list()
    prec = _lib.precision_id(args.precision)
    packed = model._packed.get(model._cfg, params) if prec == _lib.PRECISION_BF16 else None

    c2w, pix, z, tgt, sc, tc = synthetic_batch(n_obj, rank)
    n_rays = n_obj * RAYS_PER_OBJECT
    # device-resident copies (the `value` leg) and pinned host copies (the `e2e` leg)
    d_c2w, d_pix, d_z = torch.from_numpy(c2w).to(dev), torch.from_numpy(pix).to(dev), torch.from_numpy(z).to(dev)
    d_tgt, d_sc, d_tc = torch.from_numpy(tgt).to(dev), torch.from_numpy(sc).to(dev), torch.from_numpy(tc).to(dev)
    h_c2w, h_z, h_tgt = (torch.from_numpy(a).pin_memory() for a in (c2w, z, tgt))
    e_c2w, e_z, e_tgt = torch.empty_like(d_c2w), torch.empty_like(d_z), torch.empty_like(d_tgt)
    h_loss = torch.empty(n_obj, dtype=torch.float32).pin_memory()
    focal = torch.tensor([syn.SRN_FOCAL], dtype=torch.float64)
    dP = torch.zeros(sum(p.numel() for p in params), device=dev)

    def make_bundle(c2w_t, z_t):
        return cn.RayBundle(z_vals=z_t, rays_per_segment=RAYS_PER_OBJECT, c2w=c2w_t, pix_begin=d_pix, focal=focal,
                            H=syn.SRN_HW, W=syn.SRN_HW)

    def step(c2w_t, z_t, tgt_t):
        rb = make_bundle(c2w_t, z_t).args(d_sc, d_tc)
        dP.zero_()
        if prec == _lib.PRECISION_BF16:      # a training step follows an optimiser step: the bf16 operand copies are rebuilt
            model._packed.get(model._cfg, params, refresh=True)
        out = ops.render_train_step(model._cfg, params, packed, rb, prec, tgt_t, 1.0, dP, want_outputs=False)
        if world > 1:
            dist.all_reduce(dP)             # the one collective of the path: 2.86 MB MLP gradient
        return out[3]

    def step_e2e():
        e_c2w.copy_(h_c2w, non_blocking=True)
        e_z.copy_(h_z, non_blocking=True)
        e_tgt.copy_(h_tgt, non_blocking=True)
        sq = step(e_c2w, e_z, e_tgt)
        h_loss.copy_(sq, non_blocking=True)

    def fwd_only():
        rb = make_bundle(d_c2w, d_z).args(d_sc, d_tc)
        return ops.render_forward(model._cfg, params, packed, rb, prec)

    def timed(fn, steps, warmup, sample_clocks=False):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        sampler = None
        if sample_clocks and rank == 0:
            sampler = ClockSampler(local)
            sampler.start()
            time.sleep(0.25)
        L.cnb_profile_enable(1)
        l0 = L.cnb_launch_count()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()              # every rank enters the timed region together
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        launches = L.cnb_launch_count() - l0
        kt = {}
        buf = (ctypes.c_float * 512)()
        for kid, name in ((0, "fwd"), (1, "bwd"), (2, "wgrad")):
            n = L.cnb_profile_read(kid, buf, 512)
            if n > 0:
                kt[name] = [float(buf[i]) for i in range(n)]
        L.cnb_profile_enable(0)
        clocks = sampler.finish() if sampler else None
        return float(ms.item()) / steps, launches, kt, clocks

    d_seed = torch.from_numpy(tgt).to(dev) * 1e-4       # any per-ray d_rgb: the latent-fit leg times the kernels only

    def latent_only():
        # optimize.py's step body: gradients for the codes only (no weight-gradient pass, no HBM stash)
        rb = make_bundle(d_c2w, d_z).args(d_sc, d_tc)
        return ops.render_backward(model._cfg, params, packed, rb, prec, d_seed, None, False)

    ms_step, launches, kt, clocks = timed(lambda: step(d_c2w, d_z, d_tgt), args.steps, args.warmup, sample_clocks=True)
    ms_e2e, _, _, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    ms_fwd, _, kt_f, _ = timed(fwd_only, args.steps, 2)
    ms_lat, _, _, _ = timed(latent_only, args.steps, 2)
    timeouts = L.cnb_debug_pipeline_timeouts()

    if rank == 0:
        peaks = measured_peaks()
        total_rays = n_rays * world
        value = total_rays / (ms_step * 1e-3)
        e2e = total_rays / (ms_e2e * 1e-3)
        samples = n_rays * N_SAMPLES
        # dominant kernel: the one with the largest share of the step
        kmean = {k: float(np.mean(v)) * (len(v) / args.steps) for k, v in kt.items()}     # ms per step
        roof = None
        if kmean:
            dom = max(kmean, key=kmean.get)
            per_launch_ms = float(np.mean(kt[dom]))
            launches_per_step = len(kt[dom]) / args.steps
            flop_per_sample = {"fwd": FLOP_FWD, "bwd": FLOP_FWD + FLOP_DGRAD, "wgrad": FLOP_FWD}[dom]
            achieved = samples / launches_per_step * flop_per_sample / (per_launch_ms * 1e-3) / 1e12
            peak = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
            traffic = None          # dram bytes per launch of that kernel from the committed ncu --set full capture
            tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
            if os.path.exists(tpath):
                t = json.load(open(tpath))
                if t.get("objects") == n_obj and dom in t.get("dram_bytes_per_launch", {}):
                    traffic = t["dram_bytes_per_launch"][dom]
            roof = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peaks["source"] + " sustained",
                    "ms_per_launch": per_launch_ms, "share_of_step": kmean[dom] / ms_step}
        fwd_tflops = samples * FLOP_FWD / (ms_fwd * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if prec == _lib.PRECISION_BF16 else "f32", "data": "synthetic",
            "config": {"workload": f"srncar.json training step: {n_obj} objects x {RAYS_PER_OBJECT} rays x {N_SAMPLES} samples "
                                   "per GPU, fused forward + L2 loss + backward (weights, biases, codes)",
                       "net": "W=256, 3 shape + 1 texture blocks, latent 256", "view": "128x128",
                       "l2_flush": "inputs+activations far exceed L2 (per-step working set > 1 GB)",
                       "collective": "all_reduce(MLP grad 2.86 MB)" if world > 1 else "none"},
            "e2e": {"value": e2e, "unit": UNIT,
                    "h2d_bytes_per_step": int(h_c2w.numel() * 4 + h_z.numel() * 4 + h_tgt.numel() * 4),
                    "d2h_bytes_per_step": int(h_loss.numel() * 4)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "samples_per_s": value * N_SAMPLES,
            "train_tflops_algorithmic": value * N_SAMPLES * FLOP_TRAIN / 1e12,
            "train_frac_of_bf16_peak": value * N_SAMPLES * FLOP_TRAIN / 1e12 / (peaks["bf16_tflops"] * world),
            "fwd_rays_per_s": total_rays / (ms_fwd * 1e-3),
            "fwd_tflops": fwd_tflops, "fwd_frac_of_bf16_peak": fwd_tflops / peaks["bf16_tflops"],
            "latent_fit_rays_per_s": total_rays / (ms_lat * 1e-3),
            "latent_fit_frac_of_bf16_peak": samples * (FLOP_FWD + FLOP_DGRAD) / (ms_lat * 1e-3) / 1e12 / peaks["bf16_tflops"],
            "kernel_ms_per_step": kmean,
            "pipeline_timeouts": int(timeouts),
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--objects", type=int, default=32, help="objects (2048-ray segments) per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--ref-rays", type=int, default=512, help="rays per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
