"""Deterministic synthetic SRN-shaped inputs (numpy only, no torch RNG).

The dataset (ShapeNet-SRN) and trained checkpoints are not available offline, so
tests, golden fixtures and bench.py all draw their weights, codes, poses and target
pixels from the counter-based generator below; the same seed gives the same bytes
on every machine.  Shapes follow the reference: jsonfiles/srncar.json /
srnchair.json (network, near/far, N_samples), src/data.py:13-18 (pose convention),
src/trainer.py:138-139 (code init scale).
"""
import math

import numpy as np

SRN_NET = dict(shape_blocks=3, texture_blocks=1, W=256, num_xyz_freq=10, num_dir_freq=4, latent_dim=256)
SRN_CARS = dict(near=0.8, far=1.8, radius=1.3, N_samples=96)      # jsonfiles/srncar.json:15-17
SRN_CHAIRS = dict(near=1.25, far=2.75, radius=2.0, N_samples=96)  # jsonfiles/srnchair.json:15-17
SRN_FOCAL = 131.25
SRN_HW = 128


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    return z ^ (z >> np.uint64(31))


def uniform(seed, n, lo=0.0, hi=1.0):
    """n float64 uniforms in [lo, hi): hash of (seed, index)."""
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) + (np.uint64(seed) << np.uint64(40))
        h = _splitmix64(_splitmix64(idx))
    u = (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return lo + (hi - lo) * u


def normal(seed, n):
    """n float64 standard normals (Box-Muller over `uniform`)."""
    u1 = uniform(seed * 2 + 1, n)
    u2 = uniform(seed * 2 + 2, n)
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * math.pi * u2)


def param_shapes(cfg=None):
    """[(state_dict key, shape)] in CodeNeRF.state_dict() order (reference src/model.py:20-34)."""
    c = dict(SRN_NET)
    if cfg:
        c.update(cfg)
    W, LD = c["W"], c["latent_dim"]
    dx, dd = 3 + 6 * c["num_xyz_freq"], 3 + 6 * c["num_dir_freq"]
    out = [("encoding_xyz.0.weight", (W, dx)), ("encoding_xyz.0.bias", (W,))]
    for j in range(1, c["shape_blocks"] + 1):
        out += [(f"shape_latent_layer_{j}.0.weight", (W, LD)), (f"shape_latent_layer_{j}.0.bias", (W,)),
                (f"shape_layer_{j}.0.weight", (W, W)), (f"shape_layer_{j}.0.bias", (W,))]
    out += [("encoding_shape.weight", (W, W)), ("encoding_shape.bias", (W,)),
            ("sigma.0.weight", (1, W)), ("sigma.0.bias", (1,)),
            ("encoding_viewdir.0.weight", (W, W + dd)), ("encoding_viewdir.0.bias", (W,))]
    for j in range(1, c["texture_blocks"] + 1):
        out += [(f"texture_latent_layer_{j}.0.weight", (W, LD)), (f"texture_latent_layer_{j}.0.bias", (W,)),
                (f"texture_layer_{j}.0.weight", (W, W)), (f"texture_layer_{j}.0.bias", (W,))]
    out += [("rgb.0.weight", (W // 2, W)), ("rgb.0.bias", (W // 2,)),
            ("rgb.2.weight", (3, W // 2)), ("rgb.2.bias", (3,))]
    return out


def make_params(seed=0, cfg=None, gain=1.0):
    """Flat fp32 parameter vector, nn.Linear-style init: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for
    weight and bias (what kaiming_uniform_(a=sqrt(5)) gives).  Returns (flat, {key: view})."""
    shapes = param_shapes(cfg)
    total = sum(int(np.prod(s)) for _, s in shapes)
    flat = np.empty(total, np.float32)
    views, o, fan_in = {}, 0, 1
    for t, (k, shp) in enumerate(shapes):
        n = int(np.prod(shp))
        if len(shp) == 2:
            fan_in = shp[1]
        bound = gain / math.sqrt(fan_in)
        flat[o:o + n] = uniform(seed * 1000 + t, n, -bound, bound).astype(np.float32)
        views[k] = flat[o:o + n].reshape(shp)
        o += n
    return flat, views


def make_codes(seed, n, latent_dim=256):
    """randn(n, latent)/sqrt(latent/2) as reference src/trainer.py:138-139."""
    return (normal(seed, n * latent_dim) / math.sqrt(latent_dim / 2)).astype(np.float32).reshape(n, latent_dim)


def look_at_pose(seed, radius):
    """Camera on a sphere of `radius` looking at the origin, in the reference's convention:
    OpenCV-style cam-to-world from SRN pose files multiplied by diag(1,-1,-1,1) (src/data.py:13-18),
    i.e. camera looks along -z, y up.  Returns fp32 [4,4]."""
    u = uniform(seed + 7919, 2)
    theta = 2.0 * math.pi * u[0]
    zc = 0.1 + 0.8 * u[1]                       # upper hemisphere, like SRN cars/chairs
    r = math.sqrt(max(0.0, 1.0 - zc * zc))
    eye = radius * np.array([r * math.cos(theta), r * math.sin(theta), zc])
    fwd = -eye / np.linalg.norm(eye)            # viewing direction
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    true_up = np.cross(right, fwd)
    c2w = np.eye(4)
    c2w[:3, 0] = right
    c2w[:3, 1] = true_up
    c2w[:3, 2] = -fwd                           # camera -z is the viewing direction
    c2w[:3, 3] = eye
    return c2w.astype(np.float32)


def make_targets(seed, n_rays):
    """Target pixels in [0,1): fp32 [n_rays, 3] (images are /255 floats in src/data.py:37)."""
    return uniform(seed + 104729, n_rays * 3).astype(np.float32).reshape(n_rays, 3)
