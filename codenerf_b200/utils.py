"""Drop-in for the hot-path functions of the reference's src/utils.py: get_rays,
sample_from_rays, volume_rendering -- same signatures and return arity, CUDA tensors out."""
import torch

from . import _lib, ops


def _device_of(*tensors):
    _lib.require_cuda()
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    return torch.device("cuda", torch.cuda.current_device())


def _focal_args(focal):
    """The reference divides by `focal` under torch type promotion (src/utils.py:14): an fp64
    tensor with >= 1 dim (what DataLoader collation yields, src/data.py:34) promotes the pixel
    offsets to fp64; a python float or 0-dim tensor keeps fp32."""
    if isinstance(focal, torch.Tensor):
        is_f64 = focal.dtype == torch.float64 and focal.dim() >= 1
        if focal.numel() != 1:
            raise ValueError("focal must hold one value")
        val = float(focal.reshape(-1)[0].item())
        if not is_f64:
            val = float(torch.tensor(val, dtype=torch.float32).item()) if focal.dtype != torch.float64 else val
        return val, is_f64
    return float(focal), False


def collated_focal(focal):
    """What the reference's loops hand to get_rays: `data.SRN.__getitem__` returns a python float (src/data.py:34) and
    DataLoader collation turns it into an fp64 tensor of shape [1], so the pixel offsets are divided in fp64
    (`_focal_args`).  The loop mirrors (Trainer, CodeFitter) pass every python-float focal through this."""
    if isinstance(focal, torch.Tensor):
        return focal
    return torch.tensor([float(focal)], dtype=torch.float64)


def get_rays(H, W, focal, c2w):
    """Reference src/utils.py:10-19 -> (rays_o [H*W,3], viewdirs [H*W,3]), bit-exact with the
    reference's CPU result.  c2w: [4,4] (or [...,4,4] with one matrix)."""
    dev = _device_of(c2w)
    f, is_f64 = _focal_args(focal)
    c2w = torch.as_tensor(c2w)
    return ops.get_rays(int(H), int(W), f, is_f64, ops._f32c(c2w, dev))


def make_z_vals(near, far, N_samples, z_fixed=False):
    """z_vals of reference src/utils.py:24-29, evaluated with the reference's own host arithmetic:
    N scalars from python floats and the CPU generator (so `torch.manual_seed` reproduces the
    reference's jitter bit for bit).  Returns a CPU fp32 tensor [N]."""
    if z_fixed:
        return torch.linspace(near, far, N_samples)
    dist = (far - near) / (2 * N_samples)
    z_vals = torch.linspace(near + dist, far - dist, N_samples)
    z_vals += torch.rand(N_samples) * (far - near) / (2 * N_samples)
    return z_vals


def sample_from_rays(ro, vd, near, far, N_samples, z_fixed=False):
    """Reference src/utils.py:21-32 -> (xyz [R,N,3], viewdir [R,N,3], z_vals [N])."""
    dev = _device_of(ro, vd)
    z = make_z_vals(near, far, N_samples, z_fixed).to(dev)
    ro, vd = ops._f32c(ro, dev), ops._f32c(vd, dev)
    xyz, vdr = ops.sample_from_rays(ro, vd, z)
    return xyz, vdr, z


class _VolumeRendering(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sigmas, rgbs, z_vals, white_bg):
        rgb, depth, acc = ops.volume_rendering_forward(sigmas, rgbs, z_vals, white_bg)
        ctx.save_for_backward(sigmas, rgbs, z_vals)
        ctx.white_bg = white_bg
        ctx.mark_non_differentiable(acc)
        return rgb, depth, acc

    @staticmethod
    def backward(ctx, d_rgb, d_depth, _d_acc):
        sigmas, rgbs, z_vals = ctx.saved_tensors
        if d_rgb is None:
            d_rgb = torch.zeros(sigmas.numel() // z_vals.numel(), 3, device=sigmas.device)
        ds, dc = ops.volume_rendering_backward(sigmas, rgbs, z_vals, ctx.white_bg, ops._f32c(d_rgb),
                                               ops._f32c(d_depth) if d_depth is not None else None)
        return ds.view_as(sigmas), dc.view_as(rgbs), None, None


def volume_rendering_with_acc(sigmas, rgbs, z_vals, white_bg=True):
    """volume_rendering plus the accumulation map weights.sum (reference src/utils.py:45)."""
    dev = _device_of(sigmas, rgbs)
    z = ops._f32c(torch.as_tensor(z_vals), dev)
    if z.dim() != 1:
        raise ValueError("z_vals must be 1-D [N] (shared by all rays), as in the reference")
    N = z.numel()
    s, c = ops._f32c(sigmas, dev), ops._f32c(rgbs, dev)
    if s.numel() % N != 0 or c.numel() != 3 * s.numel():
        raise ValueError("sigmas must be [B,N,1] and rgbs [B,N,3]")
    return _VolumeRendering.apply(s, c, z, bool(white_bg))


def volume_rendering(sigmas, rgbs, z_vals, white_bg=True):
    """Reference src/utils.py:34-47 -> (rgb_final [B,3], depth_final [B])."""
    rgb, depth, _ = volume_rendering_with_acc(sigmas, rgbs, z_vals, white_bg)
    return rgb, depth
