"""Fused render path: get_rays -> sample_from_rays -> CodeNeRF -> volume_rendering as one
operator (the 4-call idiom of reference src/trainer.py:65-74 and src/optimizer.py:75-83,
:113-121), with autograd.  Densities and colours never reach HBM in the bf16 forward."""
import torch

from . import _lib, ops
from .utils import _focal_args


class RayBundle:
    """A batch of equal segments of rays; see cnb_ray_batch in include/codenerf_b200.h.

    Either `rays_o`/`viewdirs` ([R,3]) or cameras (`c2w` [n_seg,4,4] + focal/H/W [+ pix_begin])."""

    def __init__(self, *, z_vals, rays_per_segment, n_rays=None, rays_o=None, viewdirs=None, c2w=None,
                 pix_begin=None, focal=None, H=0, W=0, segments_per_code=1, white_bg=True):
        self.z_vals = z_vals
        self.N = z_vals.shape[-1]
        self.rays_per_segment = int(rays_per_segment)
        self.rays_o, self.viewdirs, self.c2w, self.pix_begin = rays_o, viewdirs, c2w, pix_begin
        if rays_o is not None:
            n_rays = rays_o.shape[0]
        elif n_rays is None:
            n_rays = (c2w.numel() // 16) * self.rays_per_segment
        self.n_rays = int(n_rays)
        self.focal, self.focal_is_f64 = _focal_args(focal) if focal is not None else (0.0, False)
        self.H, self.W = int(H), int(W)
        self.segments_per_code = int(segments_per_code)
        self.white_bg = bool(white_bg)

    def args(self, shape_codes, tex_codes):
        return ops.RayBatchArgs(n_rays=self.n_rays, rays_per_segment=self.rays_per_segment, n_samples=self.N,
                                z_vals=self.z_vals, shape_codes=shape_codes, tex_codes=tex_codes,
                                white_bg=self.white_bg, rays_o=self.rays_o, viewdirs=self.viewdirs, c2w=self.c2w,
                                pix_begin=self.pix_begin, focal=self.focal, focal_is_f64=self.focal_is_f64,
                                H=self.H, W=self.W, segments_per_code=self.segments_per_code)


class _RenderFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, bundle, shape_codes, tex_codes, *params):
        cfg, prec = module._cfg, _lib.precision_id(module.precision)
        training = any(p.requires_grad for p in params)         # see ops.PackedWeights: versions are not reliable then
        packed = module._packed.get(cfg, params, refresh=training) if prec == _lib.PRECISION_BF16 else None
        rb = bundle.args(shape_codes, tex_codes)
        rgb, depth, acc = ops.render_forward(cfg, params, packed, rb, prec)
        ctx.module, ctx.bundle = module, bundle
        ctx.save_for_backward(shape_codes, tex_codes, *params)
        ctx.mark_non_differentiable(acc)
        return rgb, depth, acc

    @staticmethod
    def backward(ctx, d_rgb, d_depth, _d_acc):
        module, bundle = ctx.module, ctx.bundle
        shape_codes, tex_codes, *params = ctx.saved_tensors
        cfg, prec = module._cfg, _lib.precision_id(module.precision)
        packed = module._packed.get(cfg, params) if prec == _lib.PRECISION_BF16 else None
        rb = bundle.args(shape_codes, tex_codes)
        want_p = any(ctx.needs_input_grad[4:])
        if d_rgb is None:
            d_rgb = torch.zeros(bundle.n_rays, 3, device=rb.device)
        dP, dsc, dtc = ops.render_backward(cfg, params, packed, rb, prec, ops._f32c(d_rgb),
                                           ops._f32c(d_depth) if d_depth is not None else None, want_p)
        grads = ops.split_flat_grads(cfg, dP, params) if want_p else [None] * len(params)
        return (None, None, dsc.view_as(shape_codes) if ctx.needs_input_grad[2] else None,
                dtc.view_as(tex_codes) if ctx.needs_input_grad[3] else None, *grads)


def render(model, bundle, shape_codes, tex_codes):
    """Fused render with autograd -> (rgb [R,3], depth [R], acc [R]).
    shape_codes / tex_codes: [n_codes, latent]; n_codes == 1 broadcasts, otherwise code row
    g // segments_per_code serves segment g."""
    _lib.require_cuda()
    params = model.param_list()
    dev = params[0].device
    sc = shape_codes.reshape(-1, model.latent_dim).to(dev).float().contiguous()
    tc = tex_codes.reshape(-1, model.latent_dim).to(dev).float().contiguous()
    return _RenderFunction.apply(model, bundle, sc, tc, *params)


def render_view(model, H, W, focal, c2w, z_vals, shape_code, tex_code, white_bg=True, ray_begin=0, ray_count=None):
    """One view (or a pixel window of it), rays generated in-kernel: the fused equivalent of
    get_rays + sample_from_rays + model + volume_rendering for pixels [ray_begin, ray_begin+ray_count)."""
    params = model.param_list()
    dev = params[0].device
    if ray_count is None:
        ray_count = H * W - ray_begin
    pix = torch.tensor([ray_begin], dtype=torch.int32, device=dev)
    bundle = RayBundle(z_vals=z_vals.to(dev), rays_per_segment=ray_count, n_rays=ray_count,
                       c2w=torch.as_tensor(c2w).reshape(1, 4, 4).to(dev), pix_begin=pix, focal=focal, H=H, W=W,
                       white_bg=white_bg)
    return render(model, bundle, shape_code, tex_code)
