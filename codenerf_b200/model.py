"""Drop-in for the reference's src/model.py: PE and the CodeNeRF nn.Module.

Same constructor keywords, same state_dict keys and shapes (so the reference's models.pth
checkpoints load with load_state_dict -- src/trainer.py:166, src/optimizer.py:213), same
forward(xyz, viewdir, shape_latent, texture_latent) -> (sigmas[...,1], rgbs[...,3]) signature.
The arithmetic runs in libcodenerf_b200.so (hand-written sm_100a CUDA); there is no PyTorch
or CPU fallback.
"""
import torch
import torch.nn as nn

from . import _lib, ops


def PE(x, degree):
    """Reference src/model.py:4-7: [x, sin(2^i x) (all i), cos(2^i x) (all i)].
    Host-side helper kept for API parity (the kernels encode in registers)."""
    y = torch.cat([2. ** i * x for i in range(degree)], -1)
    return torch.cat([x, torch.sin(y), torch.cos(y)], -1)


class _MLPFunction(torch.autograd.Function):
    """CodeNeRF.forward with autograd: forward = cnb_mlp_forward, backward = cnb_mlp_backward
    (recompute; no activations are kept between the two)."""

    @staticmethod
    def forward(ctx, module, xyz, viewdir, shape_codes, tex_codes, samples_per_code, *params):
        cfg, prec = module._cfg, _lib.precision_id(module.precision)
        training = any(p.requires_grad for p in params)         # see ops.PackedWeights: versions are not reliable then
        packed = module._packed.get(cfg, params, refresh=training) if prec == _lib.PRECISION_BF16 else None
        sig, col = ops.mlp_forward(cfg, params, packed, xyz, viewdir, shape_codes, tex_codes, samples_per_code, prec)
        ctx.module, ctx.spc = module, samples_per_code
        ctx.save_for_backward(xyz, viewdir, shape_codes, tex_codes, *params)
        return sig, col

    @staticmethod
    def backward(ctx, d_sig, d_col):
        module = ctx.module
        xyz, viewdir, shape_codes, tex_codes, *params = ctx.saved_tensors
        cfg, prec = module._cfg, _lib.precision_id(module.precision)
        packed = module._packed.get(cfg, params) if prec == _lib.PRECISION_BF16 else None
        want_p = any(ctx.needs_input_grad[6:])
        d_sig = ops._f32c(d_sig) if d_sig is not None else torch.zeros(xyz.numel() // 3, device=xyz.device)
        d_col = ops._f32c(d_col) if d_col is not None else torch.zeros(xyz.numel() // 3, 3, device=xyz.device)
        dP, dsc, dtc = ops.mlp_backward(cfg, params, packed, xyz, viewdir, shape_codes, tex_codes, ctx.spc, prec,
                                        d_sig, d_col, want_p)
        grads = ops.split_flat_grads(cfg, dP, params) if want_p else [None] * len(params)
        return (None, None, None, dsc if ctx.needs_input_grad[3] else None,
                dtc if ctx.needs_input_grad[4] else None, None, *grads)


class CodeNeRF(nn.Module):
    """Reference src/model.py:10-53.  `precision`: 'bf16' (tcgen05 tensor cores, fp32 accumulate)
    or 'fp32' (CUDA cores)."""

    def __init__(self, shape_blocks=2, texture_blocks=1, W=256,
                 num_xyz_freq=10, num_dir_freq=4, latent_dim=256, precision="bf16"):
        super().__init__()
        self.shape_blocks = shape_blocks
        self.texture_blocks = texture_blocks
        self.num_xyz_freq = num_xyz_freq
        self.num_dir_freq = num_dir_freq
        self.W, self.latent_dim = W, latent_dim
        self.precision = precision
        d_xyz, d_viewdir = 3 + 6 * num_xyz_freq, 3 + 6 * num_dir_freq
        # Parameter containers: identical module tree / state_dict keys to the reference
        # (src/model.py:20-34).  The nn.ReLU / nn.Softplus members hold no state; they are kept
        # so `encoding_xyz.0.weight` etc. keep their names.  forward() never calls them.
        self.encoding_xyz = nn.Sequential(nn.Linear(d_xyz, W), nn.ReLU())
        for j in range(shape_blocks):
            setattr(self, f"shape_latent_layer_{j+1}", nn.Sequential(nn.Linear(latent_dim, W), nn.ReLU()))
            setattr(self, f"shape_layer_{j+1}", nn.Sequential(nn.Linear(W, W), nn.ReLU()))
        self.encoding_shape = nn.Linear(W, W)
        self.sigma = nn.Sequential(nn.Linear(W, 1), nn.Softplus())
        self.encoding_viewdir = nn.Sequential(nn.Linear(W + d_viewdir, W), nn.ReLU())
        for j in range(texture_blocks):
            setattr(self, f"texture_latent_layer_{j+1}", nn.Sequential(nn.Linear(latent_dim, W), nn.ReLU()))
            setattr(self, f"texture_layer_{j+1}", nn.Sequential(nn.Linear(W, W), nn.ReLU()))
        self.rgb = nn.Sequential(nn.Linear(W, W // 2), nn.ReLU(), nn.Linear(W // 2, 3))
        self._cfg = ops.net_config(shape_blocks=shape_blocks, texture_blocks=texture_blocks, W=W,
                                   num_xyz_freq=num_xyz_freq, num_dir_freq=num_dir_freq, latent_dim=latent_dim)
        self._packed = ops.PackedWeights()

    def param_list(self):
        """Parameters in state_dict order (the order the C ABI expects)."""
        return list(self.parameters())

    def forward(self, xyz, viewdir, shape_latent, texture_latent):
        """xyz, viewdir: [..., 3] (typically [B,N,3]); codes: [1,latent] (broadcast, src/trainer.py:70)
        or [B,1,latent] with xyz [B,N,3] (one code per ray).  Returns sigmas [...,1], rgbs [...,3]."""
        _lib.require_cuda()
        params = self.param_list()
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("CodeNeRF parameters are not on a CUDA device; call .to('cuda') (no CPU fallback)")
        lead = xyz.shape[:-1]
        xyz_c, vd_c = ops._f32c(xyz, dev), ops._f32c(viewdir.expand_as(xyz), dev)
        S = xyz_c.numel() // 3
        sc = shape_latent.to(dev).float()
        tc = texture_latent.to(dev).float()
        n_sc, n_tc = sc.numel() // self.latent_dim, tc.numel() // self.latent_dim
        if n_sc != n_tc:
            n = max(n_sc, n_tc)
            sc = sc.reshape(-1, self.latent_dim).expand(n, -1) if n_sc == 1 else sc
            tc = tc.reshape(-1, self.latent_dim).expand(n, -1) if n_tc == 1 else tc
            n_sc = n_tc = n
        sc2, tc2 = sc.reshape(n_sc, self.latent_dim).contiguous(), tc.reshape(n_tc, self.latent_dim).contiguous()
        if n_sc == 1:
            spc = 0
        else:
            # codes [B,1,L] broadcast against [B,N,3]: row b serves the N samples of ray b
            if S % n_sc != 0:
                raise ValueError("number of code rows must divide the number of samples")
            spc = S // n_sc
        sig, col = _MLPFunction.apply(self, xyz_c.reshape(S, 3), vd_c.reshape(S, 3), sc2, tc2, spc, *params)
        return sig.reshape(*lead, 1), col.reshape(*lead, 3)
