"""Training-step semantics of the reference's Trainer on the fused render op.

Mirrors reference src/trainer.py:34-141 for one object view: AdamW over the MLP and BOTH code
tables (lr from the step schedule, halved every `interval` iterations, weight decay 0.01 default),
codes initialised randn / sqrt(latent/2), L2 loss per 2048-ray chunk (mean over the chunk,
trainer.py:75), code-norm regulariser on the first chunk only (trainer.py:76-79), gradients
accumulated over the chunks of a view, and the reference's quirk that `zero_grad` runs inside the
view loop so only the LAST view's gradients reach `step()` (trainer.py:61-64).

The dataset reader, TensorBoard logging and checkpoint cadence are out of scope; `train_view`
takes the tensors a DataLoader batch would hold.
"""
import math

import torch
import torch.nn as nn

from . import _lib, ops
from .model import CodeNeRF
from .render import RayBundle
from .utils import make_z_vals


class Trainer:
    def __init__(self, hpams, n_objects, device="cuda", batch_size=2048, precision="bf16"):
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
        self.set_optimizers()

    def get_learning_rate(self):                                                                 # trainer.py:122-128
        model_lr, latent_lr = self.hpams["lr_schedule"][0], self.hpams["lr_schedule"][1]
        lr1 = model_lr["lr"] * 2 ** (-(self.niter // model_lr["interval"]))
        lr2 = latent_lr["lr"] * 2 ** (-(self.niter // latent_lr["interval"]))
        return lr1, lr2

    def set_optimizers(self):                                                                    # trainer.py:114-120
        lr1, lr2 = self.get_learning_rate()
        # one fused multi-tensor kernel on the GPU (same update rule; an iteration on one view is launch bound)
        self.opts = torch.optim.AdamW([
            {"params": self.model.parameters(), "lr": lr1},
            {"params": self.shape_codes.parameters(), "lr": lr2},
            {"params": self.texture_codes.parameters(), "lr": lr2}], fused=self.device.type == "cuda")

    def train_view(self, focal, H, W, imgs, poses, obj_idx):
        """One iteration of trainer.py:57-96 for one object: imgs [n_views, H*W, 3], poses [n_views, 4, 4].
        Returns the mean per-chunk L2 loss of the last view (what the reference logs as PSNR)."""
        dev, B = self.device, self.B
        n_rays = H * W
        if n_rays % B != 0:
            raise ValueError("H*W must be a multiple of the ray batch size (SRN views: 16384 or 4096 rays)")
        n_chunks = n_rays // B
        cfg, params = self.model._cfg, self.model.param_list()
        prec = _lib.precision_id(self.model.precision)
        obj_idx = int(obj_idx)
        loss_mean = None
        self.opts.zero_grad()
        for k in range(imgs.shape[0]):
            self.opts.zero_grad()                                        # trainer.py:64 (discards earlier views)
            z = make_z_vals(self.hpams["near"], self.hpams["far"], self.hpams["N_samples"]).to(dev)   # one draw per view
            # repacked once per iteration: the (fused) optimiser step does not bump the parameters' version counters
            packed = self.model._packed.get(cfg, params, refresh=(k == 0)) if prec == _lib.PRECISION_BF16 else None
            sc = self.shape_codes.weight[obj_idx:obj_idx + 1]
            tc = self.texture_codes.weight[obj_idx:obj_idx + 1]
            pix = torch.arange(n_chunks, dtype=torch.int32, device=dev) * B
            bundle = RayBundle(z_vals=z, rays_per_segment=B, n_rays=n_rays,
                               c2w=poses[k].to(dev).float().reshape(1, 4, 4).expand(n_chunks, 4, 4).contiguous(),
                               pix_begin=pix, focal=focal, H=H, W=W, segments_per_code=n_chunks)
            rb = bundle.args(sc.detach(), tc.detach())
            dP = torch.zeros(sum(p.numel() for p in params), device=dev)
            tgt = imgs[k].to(dev).float().reshape(n_rays, 3).contiguous()
            _, _, _, sq, dsc, dtc = ops.render_train_step(cfg, params, packed, rb, prec, tgt, 1.0, dP, want_outputs=False)
            for p, g in zip(params, ops.split_flat_grads(cfg, dP, params)):
                p.grad = g if p.grad is None else p.grad + g          # views of this view's flat gradient vector
            # regulariser on the first chunk (trainer.py:76-79): coef * mean(|shape| + |tex|)
            coef = self.hpams["loss_reg_coef"]
            gs = torch.zeros_like(self.shape_codes.weight)
            gt = torch.zeros_like(self.texture_codes.weight)
            gs[obj_idx] = dsc[0] + coef * sc[0].detach() / sc[0].detach().norm()
            gt[obj_idx] = dtc[0] + coef * tc[0].detach() / tc[0].detach().norm()
            self.shape_codes.weight.grad, self.texture_codes.weight.grad = gs, gt
            loss_mean = (sq / (3.0 * B)).mean()
        self.opts.step()                                                 # trainer.py:85
        self.niter += 1
        return loss_mean

    def state(self):
        """The dict reference save_models writes to models.pth (trainer.py:165-174)."""
        from .checkpoint import models_dict
        return models_dict(self.model, self.shape_codes, self.texture_codes, self.niter, self.nepoch)

    def save_models(self, save_dir, iteration=None):
        """trainer.py:165-174: `models.pth` (+ `<iteration>.pth`) in the reference's wire format."""
        from .checkpoint import save_models
        return save_models(save_dir, self.model, self.shape_codes, self.texture_codes, self.niter, self.nepoch, iteration)
