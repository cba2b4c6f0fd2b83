"""Test-time latent-code fitting and novel-view evaluation on the fused render op.

Mirrors reference src/optimizer.py:48-135: codes start from the mean of the trained code tables
(optimizer.py:215-216), AdamW on the two code vectors only (lr 1e-2, re-created with lr/2 every
`lr_half_interval` steps, optimizer.py:104-105, :192-203), gradients accumulated over all target
views of a step, per-chunk mean L2 + the code-norm regulariser on each view's first chunk, then a
no-grad render of the held-out views with the jitter still on (optimizer.py:108-124) and
PSNR = -10 log10(mean of per-chunk MSEs) (optimizer.py:177-182).  The model weights take no
gradient here: the backward runs without the weight-gradient pass or any HBM stash.

`CodeFitter.fit` is the reference's loop for one object.  `fit_batch` / `evaluate_batch` /
`render_dataset` are the same arithmetic organised for the hardware: MANY test objects (or
(object, view) pairs) per fused launch with per-segment codes, per-object AdamW state (AdamW is
elementwise, so one optimiser over the [n_obj, latent] table IS n_obj independent optimisers), and
the objects / pairs sharded over the ranks of a process group with no collective on the path
(`parallel.shard_range`; results are gathered once at the end, host-side plumbing).
"""
import math

import torch

from . import _lib, ops, parallel
from .render import RayBundle, render
from .utils import collated_focal, make_z_vals


def psnr_from_chunk_mse(mse_per_chunk):
    return -10.0 * math.log(float(torch.as_tensor(mse_per_chunk).mean())) / math.log(10.0)


def _view_segments(poses, z_rows, B, n_rays, focal, H, W, dev):
    """RayBundle of whole views: poses [n, 4, 4], z_rows [n, N] -> n * n_chunks segments of B rays, one code row
    per view (segments_per_code = n_chunks)."""
    n = poses.shape[0]
    n_chunks = n_rays // B
    N = z_rows.shape[-1]
    c2w = poses.to(dev, non_blocking=True).float().reshape(n, 1, 4, 4).expand(n, n_chunks, 4, 4)
    z = z_rows.to(dev, non_blocking=True).float().reshape(n, 1, N).expand(n, n_chunks, N)
    pix = (torch.arange(n_chunks, dtype=torch.int32, device=dev) * B).repeat(n)
    return RayBundle(z_vals=z.reshape(n * n_chunks, N), rays_per_segment=B, n_rays=n * n_rays,
                     c2w=c2w.reshape(n * n_chunks, 4, 4), pix_begin=pix, focal=focal, H=H, W=W,
                     segments_per_code=n_chunks)


class CodeFitter:
    def __init__(self, model, hpams, batch_size=2048, num_opts=200):
        self.model, self.hpams, self.B, self.num_opts = model, hpams, batch_size, num_opts
        for p in self.model.parameters():          # reference never steps them (optimizer.py:195-198)
            p.requires_grad_(False)

    @classmethod
    def from_checkpoint(cls, path, hpams, device="cuda", precision="bf16", **kw):
        """optimizer.py:205-216: build the module, load `models.pth` on the CPU, move it to the device; returns
        (fitter, mean shape code, mean texture code) -- the initial codes of every test object."""
        from .checkpoint import load_models
        from .model import CodeNeRF
        model = CodeNeRF(**hpams["net_hyperparams"], precision=precision)
        _, mean_shape, mean_texture = load_models(path, model)
        return cls(model.to(device), hpams, **kw), mean_shape, mean_texture

    def _bundle(self, focal, H, W, pose, z, dev):
        n_rays = H * W
        n_chunks = n_rays // self.B
        pix = torch.arange(n_chunks, dtype=torch.int32, device=dev) * self.B
        return RayBundle(z_vals=z, rays_per_segment=self.B, n_rays=n_rays,
                         c2w=pose.to(dev).float().reshape(1, 4, 4).expand(n_chunks, 4, 4).contiguous(),
                         pix_begin=pix, focal=focal, H=H, W=W, segments_per_code=n_chunks)

    def _z(self):
        return make_z_vals(self.hpams["near"], self.hpams["far"], self.hpams["N_samples"])

    # ---- one object: the reference's loop ---------------------------------------------------------
    def fit(self, focal, H, W, tgt_imgs, tgt_poses, mean_shape, mean_texture, lr=1e-2, lr_half_interval=50):
        """tgt_imgs [n_views, H*W, 3]; returns (shapecode, texturecode, [psnr per step])."""
        dev = next(self.model.parameters()).device
        focal = collated_focal(focal)
        if (H * W) % self.B != 0:
            raise ValueError("H*W must be a multiple of the ray batch size")
        shapecode = mean_shape.to(dev).clone().detach().reshape(1, -1).requires_grad_()      # optimizer.py:64-65
        texturecode = mean_texture.to(dev).clone().detach().reshape(1, -1).requires_grad_()
        coef = self.hpams["loss_reg_coef"]
        nopts, history = 0, []

        def make_opt():
            cur = lr * 2 ** (-(nopts // lr_half_interval))                                       # optimizer.py:200-203
            return torch.optim.AdamW([{"params": shapecode, "lr": cur}, {"params": texturecode, "lr": cur}])

        opt = make_opt()
        while nopts < self.num_opts:
            opt.zero_grad()
            mses = None
            for v in range(tgt_imgs.shape[0]):
                z = self._z().to(dev)
                rgb, _, _ = render(self.model, self._bundle(focal, H, W, tgt_poses[v], z, dev), shapecode, texturecode)
                tgt = tgt_imgs[v].to(dev).float().reshape(-1, 3)
                mses = ((rgb - tgt) ** 2).reshape(-1, self.B * 3).mean(1)                         # per-chunk loss_l2
                reg = coef * torch.mean(torch.norm(shapecode, dim=-1) + torch.norm(texturecode, dim=-1))
                (mses.sum() + reg).backward()                                                     # optimizer.py:85-92
            opt.step()
            history.append(psnr_from_chunk_mse(mses.detach()))
            nopts += 1
            if nopts % lr_half_interval == 0:
                opt = make_opt()
        return shapecode.detach(), texturecode.detach(), history

    @torch.no_grad()
    def evaluate(self, focal, H, W, imgs, poses, shapecode, texturecode):
        """PSNR per held-out view (optimizer.py:108-125)."""
        dev = next(self.model.parameters()).device
        focal = collated_focal(focal)
        out = []
        for v in range(imgs.shape[0]):
            z = self._z().to(dev)
            rgb, _, _ = render(self.model, self._bundle(focal, H, W, poses[v], z, dev), shapecode, texturecode)
            tgt = imgs[v].to(dev).float().reshape(-1, 3)
            out.append(psnr_from_chunk_mse(((rgb - tgt) ** 2).reshape(-1, self.B * 3).mean(1)))
        return out

    # ---- many objects per launch, sharded over ranks ---------------------------------------------
    def fit_batch(self, focal, H, W, tgt_imgs, tgt_poses, mean_shape, mean_texture, lr=1e-2, lr_half_interval=50,
                  z_vals=None, gather=True):
        """The loop of `fit` for a batch of test objects: tgt_imgs [n_obj, n_views, H*W, 3], tgt_poses
        [n_obj, n_views, 4, 4].  Every step runs ONE fused launch per target view over all of this rank's objects
        (forward + per-chunk L2 + backward to the codes, `cnb_render_train_step` without the weight-gradient pass);
        objects are sharded over the ranks (`parallel.shard_range`), nothing is exchanged during the fit.

        z jitter: one row per (step, view, object) from the CPU generator, object-minor (the reference fits objects
        one after the other, so its stream cannot be reproduced by a batched loop; `z_vals`
        [num_opts, n_views, n_obj, N] overrides the draw, e.g. for parity tests).
        Returns (shape codes [n_obj, latent], texture codes, psnr history [num_opts, n_obj]) for ALL objects on
        every rank when `gather`, else for this rank's shard."""
        dev = next(self.model.parameters()).device
        focal = collated_focal(focal)
        B, n_rays = self.B, int(H) * int(W)
        if n_rays % B != 0:
            raise ValueError("H*W must be a multiple of the ray batch size")
        n_chunks = n_rays // B
        rank, ws = parallel.world()
        n_all, n_views = tgt_imgs.shape[0], tgt_imgs.shape[1]
        b, e = parallel.shard_range(n_all, ws, rank)
        n_obj = e - b
        latent = mean_shape.numel()
        cfg, params = self.model._cfg, self.model.param_list()
        prec = _lib.precision_id(self.model.precision)
        packed = self.model._packed.get(cfg, params) if prec == _lib.PRECISION_BF16 else None
        shape = mean_shape.to(dev).float().reshape(1, latent).repeat(max(n_obj, 1), 1).requires_grad_()
        tex = mean_texture.to(dev).float().reshape(1, latent).repeat(max(n_obj, 1), 1).requires_grad_()
        coef = self.hpams["loss_reg_coef"]
        N = self.hpams["N_samples"]
        tgts = [tgt_imgs[b:e, v].to(dev).float().reshape(n_obj * n_rays, 3) for v in range(n_views)]
        poses = [tgt_poses[b:e, v] for v in range(n_views)]
        nopts = 0
        history = torch.zeros(self.num_opts, max(n_obj, 1), device=dev)

        def make_opt():
            cur = lr * 2 ** (-(nopts // lr_half_interval))                                       # optimizer.py:200-203
            return torch.optim.AdamW([{"params": shape, "lr": cur}, {"params": tex, "lr": cur}],
                                     fused=dev.type == "cuda")

        opt = make_opt()
        gs, gt = torch.zeros_like(shape), torch.zeros_like(tex)
        while nopts < self.num_opts and n_obj > 0:
            gs.zero_(); gt.zero_()
            sq = None
            for v in range(n_views):
                if z_vals is not None:
                    z_rows = z_vals[nopts, v, b:e]
                else:
                    z_all = torch.stack([self._z() for _ in range(n_all)])      # every rank draws the same stream
                    z_rows = z_all[b:e]
                rb = _view_segments(poses[v], z_rows, B, n_rays, focal, H, W, dev).args(shape.detach(), tex.detach())
                _, _, _, sq, dsc, dtc = ops.render_train_step(cfg, params, packed, rb, prec, tgts[v], 1.0, None,
                                                              want_outputs=False)
                s, t = shape.detach(), tex.detach()
                # regulariser of each view's first chunk (optimizer.py:87-89), per object
                gs += dsc + coef * s / s.norm(dim=-1, keepdim=True)
                gt += dtc + coef * t / t.norm(dim=-1, keepdim=True)
            shape.grad, tex.grad = gs, gt
            opt.step()
            # PSNR of the step: mean over the chunks of the LAST view (optimizer.py:96: np.mean(loss_per_img))
            history[nopts] = -10.0 * torch.log10((sq / (3.0 * B)).reshape(n_obj, n_chunks).mean(1))
            nopts += 1
            if nopts % lr_half_interval == 0:
                opt = make_opt()
        s_out, t_out, h_out = shape.detach()[:n_obj], tex.detach()[:n_obj], history[:, :n_obj]
        if gather and ws > 1:
            s_out, t_out = parallel.gather_varlen(s_out), parallel.gather_varlen(t_out)
            h_out = parallel.gather_varlen(h_out.t().contiguous()).t()
        return s_out, t_out, h_out

    @torch.no_grad()
    def evaluate_batch(self, focal, H, W, imgs, poses, shapecodes, texturecodes, views_per_launch=8, gather=True):
        """optimizer.py:108-125 for many objects: imgs [n_obj, n_views, H*W, 3], poses [n_obj, n_views, 4, 4], codes
        [n_obj, latent].  (object, view) pairs are sharded over the ranks and rendered `views_per_launch` views per
        fused launch.  Returns PSNR [n_obj, n_views]."""
        n_obj, n_views = imgs.shape[0], imgs.shape[1]
        psnr = render_dataset(self.model, self.hpams, focal, H, W, poses, shapecodes, texturecodes, targets=imgs,
                              batch_size=self.B, views_per_launch=views_per_launch, gather=gather)["psnr"]
        return psnr.reshape(n_obj, n_views) if psnr.numel() == n_obj * n_views else psnr


@torch.no_grad()
def render_dataset(model, hpams, focal, H, W, poses, shapecodes, texturecodes, targets=None, batch_size=2048,
                   views_per_launch=8, z_vals=None, keep_images=False, on_batch=None, gather=True, white_bg=True):
    """Eval-scale batch render (BASELINE config 5; the reference's loop at optimizer.py:108-130): every view of
    every object -- poses [n_obj, n_views, 4, 4], codes [n_obj, latent] -- as (object, view) pairs sharded over the
    ranks (`parallel.shard_range`, no collective) and rendered `views_per_launch` whole views per fused launch
    (>= 32 segments of `batch_size` rays at the defaults), each pair with its own code rows and its own jittered
    z row (`z_vals` [n_obj, n_views, N] overrides the CPU-generator draw).

    `targets` [n_obj, n_views, H*W, 3] -> per-pair PSNR from the mean of the per-chunk MSEs (optimizer.py:177-182).
    `on_batch(pair_indices, rgb [n, H*W, 3], depth, acc)` receives every rendered batch (e.g. to write images);
    `keep_images` returns them all (only for small sets: a view is 196 KB).
    Returns {"pairs": [n, 2] (object, view) of the returned rows, "psnr": [n] or None, "rgb": [n, H*W, 3] or None};
    rows cover ALL pairs in order when `gather`, else this rank's shard."""
    params = model.param_list()
    dev = params[0].device
    focal = collated_focal(focal)
    B, n_rays = int(batch_size), int(H) * int(W)
    if n_rays % B != 0:
        raise ValueError("H*W must be a multiple of the ray batch size")
    n_obj, n_views = poses.shape[0], poses.shape[1]
    rank, ws = parallel.world()
    b, e = parallel.shard_range(n_obj * n_views, ws, rank)
    cfg, prec = model._cfg, _lib.precision_id(model.precision)
    packed = model._packed.get(cfg, params) if prec == _lib.PRECISION_BF16 else None
    sc_all = shapecodes.to(dev).float().reshape(n_obj, -1)
    tc_all = texturecodes.to(dev).float().reshape(n_obj, -1)
    N = hpams["N_samples"]
    if z_vals is None:      # one draw per pair, pair order, the same stream on every rank
        z_vals = torch.stack([make_z_vals(hpams["near"], hpams["far"], N) for _ in range(n_obj * n_views)])
    z_flat = z_vals.reshape(n_obj * n_views, N)
    poses_flat = poses.reshape(n_obj * n_views, 4, 4)
    psnrs, images = [], []
    for p0 in range(b, e, views_per_launch):
        p1 = min(p0 + views_per_launch, e)
        idx = torch.arange(p0, p1)
        obj = idx // n_views
        bundle = _view_segments(poses_flat[p0:p1], z_flat[p0:p1], B, n_rays, focal, H, W, dev)
        bundle.white_bg = bool(white_bg)
        rb = bundle.args(sc_all[obj.to(dev)], tc_all[obj.to(dev)])
        rgb, depth, acc = ops.render_forward(cfg, params, packed, rb, prec)
        n = p1 - p0
        if targets is not None:
            tgt = targets.reshape(n_obj * n_views, n_rays, 3)[p0:p1].to(dev, non_blocking=True).float()
            mse = ((rgb.reshape(n, n_rays, 3) - tgt) ** 2).reshape(n, n_rays // B, B * 3).mean(2).mean(1)
            psnrs.append(-10.0 * torch.log10(mse))
        if on_batch is not None:
            on_batch(torch.stack([obj, idx % n_views], 1), rgb.reshape(n, n_rays, 3), depth.reshape(n, n_rays),
                     acc.reshape(n, n_rays))
        if keep_images:
            images.append(rgb.reshape(n, n_rays, 3))
    pairs = torch.stack([torch.arange(b, e) // n_views, torch.arange(b, e) % n_views], 1).to(dev)
    psnr = torch.cat(psnrs) if psnrs else (torch.zeros(0, device=dev) if targets is not None else None)
    rgb_all = torch.cat(images) if images else (torch.zeros(0, n_rays, 3, device=dev) if keep_images else None)
    if gather and ws > 1:
        pairs = parallel.gather_varlen(pairs)
        psnr = parallel.gather_varlen(psnr) if psnr is not None else None
        rgb_all = parallel.gather_varlen(rgb_all) if rgb_all is not None else None
    return {"pairs": pairs, "psnr": psnr, "rgb": rgb_all}
