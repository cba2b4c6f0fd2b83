#!/usr/bin/env python
"""bench.py -- throughput of the CodeNeRF render path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload train|fit|render] [--samples 64|96]

Default workload `train` (BASELINE.json configs[1] / [3]): srncar.json network, one training step over
`--objects` objects x 2048 rays x 64 samples per GPU: fused forward (ray gen + sampling + PE + MLP +
compositing) + L2 loss + backward (weight, bias and latent-code gradients), synthetic SRN-shaped data
(128x128 views, random-init weights and codes).  For N > 1 each rank runs the same per-GPU batch on its own
objects (weak scaling) and the flat MLP gradient is all-reduced with NCCL every step.  One JSON line is printed
by rank 0.

`value` times the step with inputs resident in HBM; `e2e` times the same step through the public ops with HOST
inputs (pinned poses / z_vals / target pixels copied in, per-segment loss copied out, every step).  Extra keys:
the full trainer iteration through `Trainer.train_batch` (host batch in, AdamW step, loss out), the N = 96 legs
(jsonfiles/srncar.json:15), forward-only render, latent-only fit, and the fit / render drivers through their
public APIs (`CodeFitter.fit_batch`, `render_dataset`).

`--workload fit` (configs[2], srnchair near/far): one latent-code optimisation step over a batch of test objects
sharded across the ranks, through `CodeFitter.fit_batch`.  `--workload render` (configs[4]): batch render of
(object, view) pairs sharded across the ranks, through `render_dataset`.

`--impl reference` and the `cpu_baseline` key time the UNMODIFIED reference (src/model.py + src/utils.py staged in
oracle/_ref/ by build(); PyTorch CPU kernels on all host cores) on a bounded sample: one whole-view ray generation
plus one 2048-ray chunk forward + backward per step; if the staged files are missing, the C/OpenMP port of oracle/.
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

UNIT = "rays/s"
RAYS_PER_OBJECT = 2048                 # reference train.py:17 batch_size
VIEW_HW = 128
FLOP_FWD = 899_328                     # per sample, SURVEY.md 8(d)
FLOP_TRAIN = 2_651_904
FLOP_DGRAD = 853_248
WGRAD_BYTES_PER_SAMPLE = 7_616         # bf16 operands K3 must read once: A 472 KB + dY 480 KB per 128-row tile
METRICS = {"train": "rays/sec (%d samples/ray) fwd+bwd training step",
           "fit": "rays/sec (%d samples/ray) fwd+bwd latent-code fit step",
           "render": "rays/sec (%d samples/ray) fwd batch render"}


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


def jittered_z(seed, n_rows, cat, n_samples):
    from codenerf_b200 import synthetic as syn
    near, far = cat["near"], cat["far"]
    dist = (far - near) / (2 * n_samples)
    z0 = np.linspace(near + dist, far - dist, n_samples, dtype=np.float32)
    jit = syn.uniform(seed, n_rows * n_samples).reshape(n_rows, n_samples).astype(np.float32)
    return (z0[None] + jit * np.float32(far - near) / np.float32(2 * n_samples)).astype(np.float32)


def synthetic_batch(n_obj, rank, n_samples=64, cat=None):
    from codenerf_b200 import synthetic as syn
    cat = cat or syn.SRN_CARS
    base = 100000 * rank
    c2w = np.stack([syn.look_at_pose(base + g, cat["radius"]) for g in range(n_obj)])
    pix = np.array([(2048 * ((base + g) % 8)) for g in range(n_obj)], np.int32)       # one of the view's 8 chunks
    z = jittered_z(base + 5, n_obj, cat, n_samples)
    tgt = syn.make_targets(base + 9, n_obj * RAYS_PER_OBJECT)
    sc, tc = syn.make_codes(base + 11, n_obj), syn.make_codes(base + 12, n_obj)
    return c2w, pix, z, tgt, sc, tc


def workload_text(args, n_obj):
    if args.workload == "train":
        return (f"srncar.json training step: {n_obj} objects x {RAYS_PER_OBJECT} rays x {args.samples} samples per GPU, "
                "fused forward + L2 loss + backward (weights, biases, codes)")
    if args.workload == "fit":
        return (f"srnchair.json latent-code optimisation step (optimize.py): {n_obj} test objects x 1 target view "
                f"(128x128 = 8 chunks of 2048 rays) x {args.samples} samples per GPU, fwd + bwd to the codes + AdamW, "
                "objects sharded over the ranks")
    return (f"eval-scale batch render: {n_obj} objects x {args.views} views x 128x128 x {args.samples} samples per GPU, "
            "(object, view) pairs sharded over the ranks")


# ---------------------------------------------------------------------------------------------
def reference_chunk(device, n_samples, backward=True):
    """The unmodified reference on one view + one chunk (oracle/ref_torch.py), or None when it is not staged."""
    from oracle import ref_torch
    if not ref_torch.available():
        return None
    import torch
    from codenerf_b200 import synthetic as syn
    flat, views = syn.make_params(0)
    c2w, pix, z, tgt, sc, tc = synthetic_batch(1, 0, n_samples)
    sd = {k: torch.from_numpy(v.copy()) for k, v in views.items()}
    return ref_torch.TrainChunk(device, n_samples, sd, sc[0], tc[0], c2w[0], tgt, net=dict(syn.SRN_NET),
                                near=syn.SRN_CARS["near"], far=syn.SRN_CARS["far"])


def reference_cpu_timing(n_samples, steps, warmup, backward=True):
    """(rays/s of a whole-view iteration, cores, kind, sample text) of the reference's CPU path on all host cores."""
    cores = os.cpu_count() or 1
    import torch
    torch.set_num_threads(cores)           # torchrun exports OMP_NUM_THREADS=1: the CPU baseline gets every core anyway
    chunk = reference_chunk("cpu", n_samples, backward)
    what = "fwd+bwd (all gradients)" if backward else "fwd"
    if chunk is not None:
        t_ray, t_chunk = chunk.timed(steps, warmup, backward)
        val = RAYS_PER_OBJECT / (t_ray / 8.0 + t_chunk)
        sample = (f"unmodified reference (src/utils.py + src/model.py, torch {torch.__version__} CPU kernels, "
                  f"{torch.get_num_threads()} threads): get_rays + sample_from_rays of one 128x128 view "
                  f"({t_ray * 1e3:.0f} ms, shared by its 8 chunks) + one 2048-ray x {n_samples}-sample chunk {what} "
                  f"({t_chunk * 1e3:.0f} ms), mean of {steps}")
        return val, torch.get_num_threads(), "reference", sample, {"t_raygen_view_ms": t_ray * 1e3, "t_chunk_ms": t_chunk * 1e3}
    from codenerf_b200 import synthetic as syn
    from oracle import oracle as orc
    orc.set_num_threads(cores)
    flat, _ = syn.make_params(0)
    c2w, pix, z, tgt, sc, tc = synthetic_batch(1, 0, n_samples)
    n_rays = RAYS_PER_OBJECT

    def step():
        fwd = orc.render(flat, 128, 128, syn.SRN_FOCAL, c2w[0], z[0], sc[:1], tc[:1], True, ray_begin=0, ray_count=n_rays)
        if backward:
            d_rgb = (2.0 * (fwd["rgb"] - tgt[:n_rays]) / (3.0 * n_rays)).astype(np.float32)
            orc.render_backward(flat, fwd, z[0], sc[:1], tc[:1], d_rgb, None, True)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    sample = (f"oracle/ C + OpenMP port of the reference ({orc.num_threads()} threads; oracle/_ref not staged): one "
              f"2048-ray x {n_samples}-sample chunk {what}, mean of {steps}")
    return n_rays / dt, orc.num_threads(), "port", sample, {"t_chunk_ms": dt * 1e3}


def run_reference(args):
    """The reference's own CPU implementation of the path on the box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    backward = args.workload != "render"
    val, cores, kind, sample, detail = reference_cpu_timing(args.samples, args.steps, args.warmup, backward)
    n_obj = args.objects
    line = {"impl": "reference", "metric": METRICS[args.workload] % args.samples, "value": val, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": RAYS_PER_OBJECT / val * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_text(args, n_obj), "net": "W=256, 3 shape + 1 texture blocks, latent 256",
                       "view": "128x128", "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, **detail},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
class Timer:
    """W untimed + K timed calls between barriers, CUDA events, max over ranks; kernel times from the library."""

    def __init__(self, torch, dist, L, dev, rank, world, local):
        self.torch, self.dist, self.L, self.dev, self.rank, self.world, self.local = torch, dist, L, dev, rank, world, local

    def __call__(self, fn, steps, warmup, sample_clocks=False):
        torch, dist, L = self.torch, self.dist, self.L
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        sampler = None
        if sample_clocks and self.rank == 0:
            sampler = ClockSampler(self.local)
            sampler.start()
            time.sleep(0.25)
        L.cnb_profile_enable(1)
        l0 = L.cnb_launch_count()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()              # every rank enters the timed region together
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
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


def roofline_entries(kt, steps, samples_per_step, peaks, clocks, n_obj, n_samples):
    """One entry per timed kernel: algorithmic FLOPs (tensor-bound kernels) or bytes (the HBM-bound weight-gradient
    GEMM) per launch / its mean CUDA-event duration, against the measured peak."""
    near_max = bool(clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and
                    clocks["sm_mhz"] >= 0.95 * clocks["sm_max_mhz"])
    burst, sust = peaks["bf16_tflops"], peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
    peak_t = burst if near_max or clocks is None else sust
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath))
        if t.get("objects") == n_obj and t.get("samples", 64) == n_samples:
            traffic = t.get("dram_bytes_per_launch", {})
    out = {}
    for name, v in kt.items():
        per_launch_ms = float(np.mean(v))
        launches_per_step = len(v) / steps
        samples_per_launch = samples_per_step / launches_per_step
        e = {"kernel": {"fwd": "k_render_fwd", "bwd": "k_mlp_bwd", "wgrad": "k_wgrad"}[name],
             "ms_per_launch": per_launch_ms, "launches_per_step": launches_per_step,
             "traffic": traffic.get(name)}
        if name == "wgrad":
            a = samples_per_launch * WGRAD_BYTES_PER_SAMPLE / (per_launch_ms * 1e-3) / 1e9
            e.update(bound="hbm", achieved=a, peak=peaks["hbm_gbs"], unit="GB/s", frac=a / peaks["hbm_gbs"],
                     peak_source=peaks["source"] + " copy bandwidth",
                     tensor_tflops=samples_per_launch * FLOP_FWD / (per_launch_ms * 1e-3) / 1e12)
        else:
            flop = FLOP_FWD if name == "fwd" else FLOP_FWD + FLOP_DGRAD
            a = samples_per_launch * flop / (per_launch_ms * 1e-3) / 1e12
            e.update(bound="tensor", achieved=a, peak=peak_t, unit="TFLOP/s", frac=a / peak_t,
                     frac_of_burst=a / burst, frac_of_sustained=a / sust,
                     peak_source=peaks["source"] + (" burst (SM clock within 5% of max during the timed region)"
                                                    if peak_t == burst else " sustained"))
        out[name] = e
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import codenerf_b200 as cn
    from codenerf_b200 import _lib, ops
    from codenerf_b200 import synthetic as syn
    from codenerf_b200.optimizer import CodeFitter, render_dataset
    from codenerf_b200.trainer import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.load()
    _lib.check(L.cnb_check_device())
    timed = Timer(torch, dist, L, dev, rank, world, local)

    N = args.samples
    n_obj = args.objects
    cfg_net = dict(syn.SRN_NET)
    flat, views = syn.make_params(0, cfg_net)
    state = {k: torch.from_numpy(v.copy()) for k, v in views.items()}
    model = cn.CodeNeRF(**cfg_net, precision=args.precision)
    model.load_state_dict(state)
    model = model.to(dev)
    params = model.param_list()
    prec = _lib.precision_id(args.precision)
    packed = model._packed.get(model._cfg, params) if prec == _lib.PRECISION_BF16 else None
    focal = torch.tensor([syn.SRN_FOCAL], dtype=torch.float64)
    peaks = measured_peaks()
    hp = {"net_hyperparams": cfg_net, "N_samples": N, "near": syn.SRN_CARS["near"], "far": syn.SRN_CARS["far"],
          "loss_reg_coef": 1e-4, "lr_schedule": [{"type": "step", "lr": 1e-4, "interval": 250000},
                                                 {"type": "step", "lr": 1e-3, "interval": 250000}]}

    # ---- API-level drivers (used by every workload) ---------------------------------------------------------
    def fit_leg(n_fit, steps, warmup, cat):
        """CodeFitter.fit_batch on n_fit objects per GPU (global set sharded inside the API), one 128x128 target view."""
        hp_fit = dict(hp, near=cat["near"], far=cat["far"])
        fm = cn.CodeNeRF(**cfg_net, precision=args.precision)
        fm.load_state_dict(state)
        fitter = CodeFitter(fm.to(dev), hp_fit, batch_size=RAYS_PER_OBJECT, num_opts=steps)
        n_all = n_fit * world
        poses = torch.from_numpy(np.stack([syn.look_at_pose(5000 + g, cat["radius"]) for g in range(n_all)])).reshape(n_all, 1, 4, 4)
        imgs = torch.from_numpy(syn.make_targets(77, n_all * VIEW_HW * VIEW_HW)).reshape(n_all, 1, VIEW_HW * VIEW_HW, 3)
        imgs = imgs.pin_memory()
        mean_s, mean_t = torch.from_numpy(syn.make_codes(21, 1)[0]), torch.from_numpy(syn.make_codes(22, 1)[0])

        def call(k):
            fitter.num_opts = k
            return fitter.fit_batch(focal, VIEW_HW, VIEW_HW, imgs, poses, mean_s, mean_t, gather=False)

        call(max(1, warmup))
        ms, launches, _, _ = timed(lambda: call(steps), 1, 0)
        return ms / steps, n_all * VIEW_HW * VIEW_HW, launches

    def render_leg(n_ren, n_views, steps, warmup, to_host):
        n_all = n_ren * world
        poses = torch.from_numpy(np.stack([[syn.look_at_pose(9000 + 300 * o + v, syn.SRN_CARS["radius"]) for v in range(n_views)]
                                           for o in range(n_all)]))
        sc, tc = torch.from_numpy(syn.make_codes(31, n_all)), torch.from_numpy(syn.make_codes(32, n_all))
        host = torch.empty(n_ren * n_views, VIEW_HW * VIEW_HW, 3).pin_memory() if to_host else None
        state_ = {"row": 0}

        def sink(pairs, rgb, depth, acc):
            n = rgb.shape[0]
            host[state_["row"]:state_["row"] + n].copy_(rgb, non_blocking=True)
            state_["row"] = (state_["row"] + n) % host.shape[0]

        def call():
            state_["row"] = 0
            render_dataset(model, hp, focal, VIEW_HW, VIEW_HW, poses, sc, tc, batch_size=RAYS_PER_OBJECT,
                           views_per_launch=8, on_batch=sink if to_host else None, gather=False)

        ms, launches, kt, clocks = timed(call, steps, warmup, sample_clocks=(args.workload == "render" and not to_host))
        return ms, n_all * n_views * VIEW_HW * VIEW_HW, launches, kt, clocks

    # ---- workload: fit / render through their public APIs ------------------------------------------------------
    if args.workload in ("fit", "render"):
        if args.workload == "fit":
            n_fit = args.objects
            ms_step, rays_step, launches = fit_leg(n_fit, args.steps, args.warmup, syn.SRN_CHAIRS)
            ms_e2e, kt, clocks = ms_step, {}, None          # the API call IS end to end: host images in, codes out
            h2d = n_fit * VIEW_HW * VIEW_HW * 3 * 4 // max(args.steps, 1) + n_fit * (64 + N * 4)
            d2h = 0
            flop = FLOP_FWD + FLOP_DGRAD
        else:
            ms_step, rays_step, launches, kt, clocks = render_leg(args.objects, args.views, args.steps, args.warmup, False)
            ms_e2e, _, _, _, _ = render_leg(args.objects, args.views, args.steps, max(1, args.warmup // 2), True)
            h2d = args.objects * args.views * (64 + N * 4)
            d2h = args.objects * args.views * VIEW_HW * VIEW_HW * 12
            flop = FLOP_FWD
        if rank == 0:
            value = rays_step / (ms_step * 1e-3)
            tfl = value * N * flop / 1e12
            roofs = roofline_entries(kt, args.steps, rays_step // world * N, peaks, clocks, args.objects, N) if kt else {}
            line = {"metric": METRICS[args.workload] % N, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "bf16" if prec == _lib.PRECISION_BF16 else "f32", "data": "synthetic",
                    "config": {"workload": workload_text(args, args.objects), "net": "W=256, 3 shape + 1 texture blocks, latent 256",
                               "view": "128x128", "l2_flush": "per-step working set far exceeds L2", "collective": "none"},
                    "e2e": {"value": rays_step / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                    "gpu_launches": int(launches), "clocks": clocks,
                    "roofline": roofs.get("fwd") or {"bound": "tensor", "kernel": "k_mlp_bwd (through CodeFitter.fit_batch)",
                                                     "achieved": tfl / world, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                                                     "frac": tfl / world / peaks["bf16_tflops"], "traffic": None,
                                                     "note": "whole API step (kernels + AdamW + host loop) over algorithmic FLOPs"},
                    "samples_per_s": value * N, "tflops_algorithmic": tfl,
                    "frac_of_bf16_peak": tfl / (peaks["bf16_tflops"] * world),
                    "pipeline_timeouts": int(L.cnb_debug_pipeline_timeouts())}
            if world == 1 and not args.no_cpu_baseline:
                v, cores, kind, sample, detail = reference_cpu_timing(N, 2, 1, backward=args.workload == "fit")
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, **detail}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- workload: train -----------------------------------------------------------------------------------
    def make_inputs(n_samples):
        c2w, pix, z, tgt, sc, tc = synthetic_batch(n_obj, rank, n_samples)
        d = dict(c2w=torch.from_numpy(c2w).to(dev), pix=torch.from_numpy(pix).to(dev), z=torch.from_numpy(z).to(dev),
                 tgt=torch.from_numpy(tgt).to(dev), sc=torch.from_numpy(sc).to(dev), tc=torch.from_numpy(tc).to(dev))
        d["h_c2w"], d["h_z"], d["h_tgt"] = (torch.from_numpy(a).pin_memory() for a in (c2w, z, tgt))
        d["e_c2w"], d["e_z"], d["e_tgt"] = torch.empty_like(d["c2w"]), torch.empty_like(d["z"]), torch.empty_like(d["tgt"])
        return d

    n_rays = n_obj * RAYS_PER_OBJECT
    dP = torch.zeros(sum(p.numel() for p in params), device=dev)
    h_loss = torch.empty(n_obj, dtype=torch.float32).pin_memory()

    def make_bundle(I, c2w_t, z_t):
        return cn.RayBundle(z_vals=z_t, rays_per_segment=RAYS_PER_OBJECT, c2w=c2w_t, pix_begin=I["pix"], focal=focal,
                            H=syn.SRN_HW, W=syn.SRN_HW)

    def step(I, c2w_t, z_t, tgt_t):
        rb = make_bundle(I, c2w_t, z_t).args(I["sc"], I["tc"])
        dP.zero_()
        if prec == _lib.PRECISION_BF16:      # a training step follows an optimiser step: the bf16 operand copies are rebuilt
            model._packed.get(model._cfg, params, refresh=True)
        out = ops.render_train_step(model._cfg, params, packed, rb, prec, tgt_t, 1.0, dP, want_outputs=False)
        if world > 1:
            dist.all_reduce(dP)             # the one collective of the path: 2.86 MB MLP gradient
        return out[3]

    def step_e2e(I):
        I["e_c2w"].copy_(I["h_c2w"], non_blocking=True)
        I["e_z"].copy_(I["h_z"], non_blocking=True)
        I["e_tgt"].copy_(I["h_tgt"], non_blocking=True)
        sq = step(I, I["e_c2w"], I["e_z"], I["e_tgt"])
        h_loss.copy_(sq, non_blocking=True)

    def fwd_only(I):
        rb = make_bundle(I, I["c2w"], I["z"]).args(I["sc"], I["tc"])
        return ops.render_forward(model._cfg, params, packed, rb, prec)

    def latent_only(I, seed):
        # optimize.py's step body: gradients for the codes only (no weight-gradient pass, no HBM stash)
        rb = make_bundle(I, I["c2w"], I["z"]).args(I["sc"], I["tc"])
        return ops.render_backward(model._cfg, params, packed, rb, prec, seed, None, False)

    I = make_inputs(N)
    ms_step, launches, kt, clocks = timed(lambda: step(I, I["c2w"], I["z"], I["tgt"]), args.steps, args.warmup, sample_clocks=True)
    ms_e2e, _, _, _ = timed(lambda: step_e2e(I), args.steps, max(1, args.warmup // 2))
    fwd_steps = max(args.steps, 20)      # a forward step is 3.5 ms: more of them for a stable number
    ms_fwd, _, kt_f, _ = timed(lambda: fwd_only(I), fwd_steps, 3)
    d_seed = I["tgt"] * 1e-4            # any per-ray d_rgb: the latent-fit leg times the kernels only
    ms_lat, _, _, _ = timed(lambda: latent_only(I, d_seed), args.steps, 2)
    extra = {}
    if N != 96 and not args.quick:      # the reference's own sample count (jsonfiles/srncar.json:15)
        I96 = make_inputs(96)
        ms96, _, kt96, _ = timed(lambda: step(I96, I96["c2w"], I96["z"], I96["tgt"]), max(2, args.steps // 2), 2)
        ms96f, _, _, _ = timed(lambda: fwd_only(I96), max(2, args.steps // 2), 2)
        tot = n_rays * world
        extra.update(n96_train_rays_per_s=tot / (ms96 * 1e-3), n96_fwd_rays_per_s=tot / (ms96f * 1e-3),
                     n96_train_tflops=tot * 96 * FLOP_TRAIN / (ms96 * 1e-3) / 1e12,
                     n96_fwd_tflops=tot * 96 * FLOP_FWD / (ms96f * 1e-3) / 1e12,
                     n96_kernel_ms_per_step={k: float(np.sum(v)) / max(2, args.steps // 2) for k, v in kt96.items()})
        del I96
    # the full trainer iteration through the public Trainer (host batch in, fused step, all-reduce, AdamW, loss out)
    if not args.quick:
        torch.manual_seed(1234 + rank)
        tr = Trainer(hp, n_objects=n_obj * world, device=dev, batch_size=RAYS_PER_OBJECT, precision=args.precision)
        tr.model.load_state_dict(state)
        c2w_np = np.stack([syn.look_at_pose(100000 * rank + g, syn.SRN_CARS["radius"]) for g in range(n_obj)])
        t_imgs = torch.from_numpy(syn.make_targets(100000 * rank + 9, n_obj * VIEW_HW * VIEW_HW)).reshape(n_obj, VIEW_HW * VIEW_HW, 3).pin_memory()
        t_poses = torch.from_numpy(c2w_np).pin_memory()
        objs = list(range(rank * n_obj, (rank + 1) * n_obj))
        # the trainer renders WHOLE views (8 chunks per object): use n_obj / 8 objects so the step has the same ray count
        n_tr = max(1, n_obj // 8)
        h_l = torch.empty(n_tr).pin_memory()

        def trainer_iter():
            loss = tr.train_batch(focal, VIEW_HW, VIEW_HW, t_imgs[:n_tr], t_poses[:n_tr], objs[:n_tr])
            h_l.copy_(loss, non_blocking=True)

        ms_tr, _, _, _ = timed(trainer_iter, args.steps, max(2, args.warmup // 2))
        extra["train_iter_e2e_rays_per_s"] = n_tr * VIEW_HW * VIEW_HW * world / (ms_tr * 1e-3)
        extra["train_iter_e2e"] = {"ms_per_iteration": ms_tr, "through": "Trainer.train_batch (host images + poses in, z jitter drawn, "
                                   "fused step, gradient all-reduce, fused AdamW over MLP + code tables, per-object loss out)",
                                   "objects_x_rays": [n_tr, VIEW_HW * VIEW_HW]}
        del tr
        ms_fit, rays_fit, _ = fit_leg(max(1, n_obj // 8), max(4, args.steps), 2, syn.SRN_CHAIRS)
        extra["fit_rays_per_s"] = rays_fit / (ms_fit * 1e-3)
        ms_ren, rays_ren, _, _, _ = render_leg(max(1, n_obj // 8), 8, max(2, args.steps // 2), 1, False)
        extra["render_rays_per_s"] = rays_ren / (ms_ren * 1e-3)
        extra["api_legs"] = {"fit": "CodeFitter.fit_batch, srnchair near/far, %d objects x 1 view per GPU" % max(1, n_obj // 8),
                             "render": "render_dataset, %d objects x 8 views per GPU" % max(1, n_obj // 8)}
    timeouts = L.cnb_debug_pipeline_timeouts()

    if rank == 0:
        total_rays = n_rays * world
        value = total_rays / (ms_step * 1e-3)
        e2e = total_rays / (ms_e2e * 1e-3)
        samples = n_rays * N
        roofs = roofline_entries(kt, args.steps, samples, peaks, clocks, n_obj, N)
        roofs.update(roofline_entries({k: v for k, v in kt_f.items() if k == "fwd"}, fwd_steps, samples, peaks, clocks, n_obj, N))
        kmean = {k: float(np.sum(v)) / args.steps for k, v in kt.items()}     # ms per step
        dom = max(kmean, key=kmean.get) if kmean else None
        roof = dict(roofs[dom], share_of_step=kmean[dom] / ms_step) if dom else None
        fwd_tflops = samples * FLOP_FWD / (ms_fwd * 1e-3) / 1e12
        line = {
            "metric": METRICS["train"] % N, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if prec == _lib.PRECISION_BF16 else "f32", "data": "synthetic",
            "config": {"workload": workload_text(args, n_obj),
                       "net": "W=256, 3 shape + 1 texture blocks, latent 256", "view": "128x128",
                       "l2_flush": "inputs+activations far exceed L2 (per-step working set > 1 GB)",
                       "collective": "all_reduce(MLP grad 2.86 MB)" if world > 1 else "none"},
            "e2e": {"value": e2e, "unit": UNIT,
                    "h2d_bytes_per_step": int(I["h_c2w"].numel() * 4 + I["h_z"].numel() * 4 + I["h_tgt"].numel() * 4),
                    "d2h_bytes_per_step": int(h_loss.numel() * 4)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "rooflines": roofs,
            "samples_per_s": value * N,
            "train_tflops_algorithmic": value * N * FLOP_TRAIN / 1e12,
            "train_frac_of_bf16_peak": value * N * FLOP_TRAIN / 1e12 / (peaks["bf16_tflops"] * world),
            "train_frac_of_bf16_sustained": value * N * FLOP_TRAIN / 1e12 / ((peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]) * world),
            "fwd_rays_per_s": total_rays / (ms_fwd * 1e-3),
            "fwd_tflops": fwd_tflops * world, "fwd_frac_of_bf16_peak": fwd_tflops / peaks["bf16_tflops"],
            "latent_fit_rays_per_s": total_rays / (ms_lat * 1e-3),
            "latent_fit_frac_of_bf16_peak": samples * (FLOP_FWD + FLOP_DGRAD) / (ms_lat * 1e-3) / 1e12 / peaks["bf16_tflops"],
            "kernel_ms_per_step": kmean,
            "pipeline_timeouts": int(timeouts),
        }
        line.update(extra)
        if world == 1 and not args.no_cpu_baseline:
            v, cores, kind, sample, detail = reference_cpu_timing(N, 2, 1, True)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, **detail}
            try:        # the reference in its own deployment mode: fp32 eager PyTorch on this GPU
                ch = reference_chunk(dev, N, True)
                if ch is not None:
                    t_ray, t_chunk = ch.timed(5, 2, True)
                    line["reference_gpu_fp32_eager"] = {
                        "rays_per_s": RAYS_PER_OBJECT / (t_ray / 8.0 + t_chunk), "chunk_only_rays_per_s": RAYS_PER_OBJECT / t_chunk,
                        "t_raygen_view_ms": t_ray * 1e3, "t_chunk_ms": t_chunk * 1e3,
                        "what": "unmodified reference functions, CodeNeRF on cuda in fp32 (TF32 off), host ray generation + per-chunk "
                                "H2D as in src/trainer.py:65-74, one 2048-ray chunk fwd+bwd"}
            except Exception as ex:            # never lose the bench line to the comparison leg
                line["reference_gpu_fp32_eager"] = {"error": repr(ex)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "fit", "render"])
    ap.add_argument("--objects", type=int, default=None, help="objects per GPU per step (train: 2048-ray segments, default 32; "
                    "fit: test objects with one 128x128 view, default 8; render: objects, default 4)")
    ap.add_argument("--views", type=int, default=16, help="render workload: views per object")
    ap.add_argument("--samples", type=int, default=64, help="depth samples per ray (BASELINE metric: 64; srncar.json: 96)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="train workload: skip the N=96, trainer and API legs")
    args = ap.parse_args()
    if args.objects is None:
        args.objects = {"train": 32, "fit": 8, "render": 4}[args.workload]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
