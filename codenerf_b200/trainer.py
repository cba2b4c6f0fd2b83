"""The reference's Trainer (src/trainer.py:15-174) on the fused render op.

What is kept from the reference, line for line in meaning:

* AdamW over the MLP and BOTH code tables, lr from the step schedule (halved every `interval`
  iterations), weight decay 0.01 (the default), re-created at the start of every epoch
  (trainer.py:52, :114-128) -- which resets the moments, `begin_epoch()` here;
* codes initialised randn / sqrt(latent/2) (trainer.py:133-141);
* per object and view: one jittered z row (trainer.py:66-67), L2 loss per 2048-ray chunk (mean over
  the chunk, :75), the code-norm regulariser on the first chunk only (:76-79), gradients accumulated
  over the chunks (:82), and the quirk that `zero_grad` runs inside the view loop (:64) so only the
  LAST view of an object reaches `step()`;
* crop -> full curriculum (`training()`, trainer.py:34-47): 64x64 centre crops with the focal
  unchanged for `iters_crop` iterations, 128x128 afterwards; `save_models` after every epoch and
  every `check_points` iterations (:94-95, :46).

What is new (B200-first): an iteration takes a BATCH of objects -- `objects_per_step`, 1 reproduces the
reference's sequence exactly -- rendered, differentiated and reduced by ONE fused launch per step
(`cnb_render_train_step`: ray generation, sampling, PE, MLP, compositing, loss, backward), and the
batch is sharded over the ranks of a process group: every rank runs its objects, the flat MLP
gradient is summed with one asynchronous NCCL all-reduce that overlaps the code-row updates, code
rows stay on the rank that owns the object (`parallel.gather_owned_rows` at checkpoints).  The step
of a batch equals the single-GPU step on the concatenated batch (gradients are summed, like the
reference's chunk accumulation).

TensorBoard / PNG logging (trainer.py:98-112) is host glue outside the path and is not mirrored.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops, parallel
from .model import CodeNeRF
from .render import RayBundle
from .utils import collated_focal, make_z_vals


class Trainer:
    def __init__(self, hpams, n_objects, device="cuda", batch_size=2048, precision="bf16", check_iter=10000):
        self.hpams = hpams
        self.device = torch.device(device)
        self.B = batch_size
        self.model = CodeNeRF(**hpams["net_hyperparams"], precision=precision).to(self.device)   # trainer.py:131
        embdim = hpams["net_hyperparams"]["latent_dim"]
        self.shape_codes = nn.Embedding(n_objects, embdim)                                       # trainer.py:133-141
        self.texture_codes = nn.Embedding(n_objects, embdim)
        self.shape_codes.weight = nn.Parameter(torch.randn(n_objects, embdim) / math.sqrt(embdim / 2))
        self.texture_codes.weight = nn.Parameter(torch.randn(n_objects, embdim) / math.sqrt(embdim / 2))
        self.shape_codes, self.texture_codes = self.shape_codes.to(self.device), self.texture_codes.to(self.device)
        self.niter, self.nepoch = 0, 0
        self.check_iter = check_iter
        self.rank, self.world = parallel.world()
        self._owner = None                      # object -> owning rank of its code rows (multi-GPU)
        n_par = sum(p.numel() for p in self.model.parameters())
        # one flat gradient vector, the parameters' .grad are views into it: the fused step accumulates into it, the
        # all-reduce and the (fused, multi-tensor) AdamW read it -- no per-iteration allocation
        self._dP = torch.zeros(n_par, device=self.device)
        self._grad_views = None
        self.set_optimizers()

    # ---- optimiser (trainer.py:114-128) --------------------------------------------------------
    def get_learning_rate(self):
        model_lr, latent_lr = self.hpams["lr_schedule"][0], self.hpams["lr_schedule"][1]
        lr1 = model_lr["lr"] * 2 ** (-(self.niter // model_lr["interval"]))
        lr2 = latent_lr["lr"] * 2 ** (-(self.niter // latent_lr["interval"]))
        return lr1, lr2

    def set_optimizers(self):
        lr1, lr2 = self.get_learning_rate()
        # one fused multi-tensor kernel on the GPU (same update rule as the reference's default AdamW)
        self.opts = torch.optim.AdamW([
            {"params": self.model.parameters(), "lr": lr1},
            {"params": self.shape_codes.parameters(), "lr": lr2},
            {"params": self.texture_codes.parameters(), "lr": lr2}], fused=self.device.type == "cuda")

    def begin_epoch(self):
        """trainer.py:52: every epoch starts with a fresh AdamW at the scheduled learning rates (moments reset)."""
        self.set_optimizers()

    # ---- one iteration ---------------------------------------------------------------------------
    def _bind_grads(self):
        params = self.model.param_list()
        if self._grad_views is None:
            self._grad_views = ops.split_flat_grads(self.model._cfg, self._dP, params)
        for p, g in zip(params, self._grad_views):
            p.grad = g
        for t in (self.shape_codes.weight, self.texture_codes.weight):
            if t.grad is None:
                t.grad = torch.zeros_like(t)
        return params

    def train_batch(self, focal, H, W, imgs, poses, obj_idx, z_vals=None):
        """One optimiser step on a batch of objects (this rank's share of the step), one fused launch.

        imgs [n_obj, n_views, H*W, 3] (or [n_obj, H*W, 3]), poses [n_obj, n_views, 4, 4] (or [n_obj, 4, 4]),
        obj_idx [n_obj] rows of the code tables.  With several views per object only the LAST one contributes, as in
        the reference (trainer.py:64); one jittered z row is drawn per object and view from the CPU generator in
        object-major order (the reference's order for objects_per_step = 1) unless `z_vals` [n_obj, N] is given.
        Returns the per-object mean of the per-chunk L2 losses of that view (device tensor [n_obj]; the reference
        logs its PSNR, trainer.py:86)."""
        dev, B = self.device, self.B
        focal = collated_focal(focal)
        n_rays = int(H) * int(W)
        if n_rays % B != 0:
            raise ValueError("H*W must be a multiple of the ray batch size (SRN views: 16384 or 4096 rays)")
        n_chunks = n_rays // B
        imgs, poses = torch.as_tensor(imgs), torch.as_tensor(poses)
        if imgs.dim() == 3:
            imgs, poses = imgs.unsqueeze(1), poses.unsqueeze(1)
        n_obj, n_views = imgs.shape[0], imgs.shape[1]
        N = self.hpams["N_samples"]
        if z_vals is None:
            rows = []
            for _ in range(n_obj):
                for k in range(n_views):                 # every view draws (RNG order), the last one is used
                    z = make_z_vals(self.hpams["near"], self.hpams["far"], N)
                rows.append(z)
            z_vals = torch.stack(rows)
        obj_idx = torch.as_tensor(obj_idx, dtype=torch.long).reshape(-1).to(dev)
        cfg = self.model._cfg
        params = self._bind_grads()
        prec = _lib.precision_id(self.model.precision)
        # repacked once per iteration: the (fused) optimiser step does not bump the parameters' version counters
        packed = self.model._packed.get(cfg, params, refresh=True) if prec == _lib.PRECISION_BF16 else None
        sc = self.shape_codes.weight.detach()[obj_idx]
        tc = self.texture_codes.weight.detach()[obj_idx]
        c2w = poses[:, -1].to(dev, non_blocking=True).float().reshape(n_obj, 1, 4, 4).expand(n_obj, n_chunks, 4, 4)
        pix = (torch.arange(n_chunks, dtype=torch.int32, device=dev) * B).repeat(n_obj)
        z_seg = z_vals.to(dev, non_blocking=True).float().reshape(n_obj, 1, N).expand(n_obj, n_chunks, N)
        bundle = RayBundle(z_vals=z_seg.reshape(n_obj * n_chunks, N), rays_per_segment=B, n_rays=n_obj * n_rays,
                           c2w=c2w.reshape(n_obj * n_chunks, 4, 4), pix_begin=pix, focal=focal, H=H, W=W,
                           segments_per_code=n_chunks)
        rb = bundle.args(sc, tc)
        tgt = imgs[:, -1].to(dev, non_blocking=True).float().reshape(n_obj * n_rays, 3)
        self._dP.zero_()
        _, _, _, sq, dsc, dtc = ops.render_train_step(cfg, params, packed, rb, prec, tgt, 1.0, self._dP,
                                                      want_outputs=False)
        # the one collective of the path: sum of the flat MLP gradient over the ranks, overlapped with the code rows
        work = parallel.allreduce_mlp_grad(self._dP, async_op=True)
        # regulariser on the first chunk (trainer.py:76-79): coef * mean(|shape| + |tex|), one code row per object
        coef = self.hpams["loss_reg_coef"]
        gs, gt = self.shape_codes.weight.grad, self.texture_codes.weight.grad
        gs.zero_(); gt.zero_()
        gs.index_add_(0, obj_idx, dsc + coef * sc / sc.norm(dim=-1, keepdim=True))
        gt.index_add_(0, obj_idx, dtc + coef * tc / tc.norm(dim=-1, keepdim=True))
        work.wait()
        self.opts.step()                                                 # trainer.py:85
        self.niter += 1
        return (sq / (3.0 * B)).reshape(n_obj, n_chunks).mean(1)

    def train_view(self, focal, H, W, imgs, poses, obj_idx):
        """One iteration of trainer.py:57-96 for ONE object: imgs [n_views, H*W, 3], poses [n_views, 4, 4].
        Returns the mean per-chunk L2 loss of the last view (what the reference logs as PSNR)."""
        loss = self.train_batch(focal, H, W, torch.as_tensor(imgs).unsqueeze(0), torch.as_tensor(poses).unsqueeze(0),
                                [int(obj_idx)])
        return loss[0]

    # ---- epochs (trainer.py:34-53) ----------------------------------------------------------------
    def training(self, dataset, iters_crop, iters_all, objects_per_step=1, save_dir=None, log=None):
        """The reference's outer loop over a `data.SRN` dataset: epochs over all objects, the crop -> full curriculum,
        a fresh optimiser per epoch, `models.pth` after every epoch and `<niter>.pth` every `check_points` iterations.
        `objects_per_step` objects are batched per iteration and sharded over the ranks; `log(niter, psnr)` is called
        with the mean PSNR of the step on rank 0."""
        if iters_crop > iters_all:
            raise OSError("iters_crop > iters_all")                       # trainer.py:35-36
        self._owner = None
        while self.niter < iters_all:
            crop = self.niter < iters_crop
            self.training_single_epoch(dataset, iters_crop if crop else iters_all, crop, objects_per_step, save_dir, log)
            if save_dir is not None:
                self.save_models(save_dir)
            self.nepoch += 1

    def training_single_epoch(self, dataset, num_iters, crop_img, objects_per_step=1, save_dir=None, log=None):
        dataset.crop_img = crop_img                                       # trainer.py:50 (make_dataloader)
        self.begin_epoch()                                                # trainer.py:52
        n = len(dataset)
        if self._owner is None:
            self._owner = torch.zeros(n, dtype=torch.int32)
        for first in range(0, n, objects_per_step):
            if self.niter >= num_iters:
                break
            batch = list(range(first, min(first + objects_per_step, n)))
            b, e = parallel.shard_range(len(batch), self.world, self.rank)
            for r in range(self.world):
                rb, re_ = parallel.shard_range(len(batch), self.world, r)
                self._owner[batch[rb:re_]] = r
            mine = batch[b:e]
            check = save_dir is not None and self.niter % self.hpams.get("check_points", 1 << 62) == 0
            if mine:
                focal, H, W, imgs, poses, _ = dataset.train_batch(mine, device=self.device)
                loss = self.train_batch(focal, H, W, imgs, poses, mine)
            else:                                       # fewer objects than ranks in the last step: gradient-less step
                loss = self._empty_step()
            if log is not None:
                tot = torch.stack([loss.sum(), torch.tensor(float(loss.numel()), device=self.device)])
                if self.world > 1:
                    torch.distributed.all_reduce(tot)
                if self.rank == 0:
                    log(self.niter - 1, -10.0 * math.log10(max(float(tot[0] / tot[1]), 1e-30)))
            if check:
                self.save_models(save_dir, self.niter - 1)                # trainer.py:94-95 (before niter += 1)

    def _empty_step(self):
        self._bind_grads()
        self._dP.zero_()
        parallel.allreduce_mlp_grad(self._dP)
        self.shape_codes.weight.grad.zero_(); self.texture_codes.weight.grad.zero_()
        self.opts.step()
        self.niter += 1
        return torch.zeros(0, device=self.device)

    # ---- checkpoints (trainer.py:165-174) -----------------------------------------------------------
    def _full_tables(self):
        """Code tables with every row taken from the rank that owns it (identity on one GPU)."""
        if self.world == 1 or self._owner is None:
            return self.shape_codes, self.texture_codes
        s = parallel.gather_owned_rows(self.shape_codes.weight, self._owner)
        t = parallel.gather_owned_rows(self.texture_codes.weight, self._owner)
        return s, t

    def state(self):
        """The dict reference save_models writes to models.pth (trainer.py:165-174)."""
        from .checkpoint import models_dict
        s, t = self._full_tables()
        return models_dict(self.model, s, t, self.niter, self.nepoch)

    def save_models(self, save_dir, iteration=None):
        """trainer.py:165-174: `models.pth` (+ `<iteration>.pth`) in the reference's wire format (rank 0 writes)."""
        from .checkpoint import save_models
        s, t = self._full_tables()
        if self.rank != 0:
            return None
        return save_models(save_dir, self.model, s, t, self.niter, self.nepoch, iteration)
