"""Test-time latent-code fitting and novel-view evaluation on the fused render op.

Mirrors reference src/optimizer.py:48-135: codes start from the mean of the trained code tables
(optimizer.py:215-216), AdamW on the two code vectors only (lr 1e-2, re-created with lr/2 every
`lr_half_interval` steps, optimizer.py:104-105, :192-203), gradients accumulated over all target
views of a step, per-chunk mean L2 + the code-norm regulariser on each view's first chunk, then a
no-grad render of the held-out views with the jitter still on (optimizer.py:108-124) and
PSNR = -10 log10(mean of per-chunk MSEs) (optimizer.py:177-182).  The model weights take no
gradient here: the backward runs without the weight-gradient pass or any HBM stash.
"""
import math

import torch

from .render import RayBundle, render
from .utils import make_z_vals


def psnr_from_chunk_mse(mse_per_chunk):
    return -10.0 * math.log(float(torch.as_tensor(mse_per_chunk).mean())) / math.log(10.0)


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

    def fit(self, focal, H, W, tgt_imgs, tgt_poses, mean_shape, mean_texture, lr=1e-2, lr_half_interval=50):
        """tgt_imgs [n_views, H*W, 3]; returns (shapecode, texturecode, [psnr per step])."""
        dev = next(self.model.parameters()).device
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
                z = make_z_vals(self.hpams["near"], self.hpams["far"], self.hpams["N_samples"]).to(dev)
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
        out = []
        for v in range(imgs.shape[0]):
            z = make_z_vals(self.hpams["near"], self.hpams["far"], self.hpams["N_samples"]).to(dev)
            rgb, _, _ = render(self.model, self._bundle(focal, H, W, poses[v], z, dev), shapecode, texturecode)
            tgt = imgs[v].to(dev).float().reshape(-1, 3)
            out.append(psnr_from_chunk_mse(((rgb - tgt) ** 2).reshape(-1, self.B * 3).mean(1)))
        return out
