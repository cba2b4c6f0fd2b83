"""Shared helpers of the GPU parity tests."""
import numpy as np
import torch

from codenerf_b200 import synthetic as syn
from oracle import oracle as orc


def make_model(precision, seed=0, cfg=None):
    import codenerf_b200 as cn
    c = dict(syn.SRN_NET)
    if cfg:
        c.update(cfg)
    flat, views = syn.make_params(seed, c)
    m = cn.CodeNeRF(**c, precision=precision)
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in views.items()})
    return m.cuda(), flat


def named_grads_flat(model):
    return np.concatenate([p.grad.detach().cpu().numpy().ravel() for p in model.parameters()])


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def cosine(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300))


def oracle_render_case(flat, H, W, focal, c2w, near, far, N, seed, shape_codes, tex_codes, white=True,
                       ray_begin=0, ray_count=None):
    z = orc.z_vals(near, far, N, orc.torch_rand(seed, N))
    fwd = orc.render(flat, H, W, focal, c2w, z, shape_codes, tex_codes, white, ray_begin=ray_begin,
                     ray_count=ray_count)
    return z, fwd
