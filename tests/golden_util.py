"""Helpers shared by the golden-fixture tests (same definitions as tests/golden/make_golden.py)."""
import os

import numpy as np

from codenerf_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def checksum(a):
    b = bits(a).ravel().astype(np.uint64)
    idx = np.arange(b.size, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return np.uint64(np.bitwise_xor.reduce((b + np.uint64(1)) * (idx * np.uint64(2654435761) + np.uint64(97))))


def weight_probe(seed, n):
    return syn.uniform(seed + 31337, n, -1.0, 1.0)


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def rays_meta():
    g = load("rays_samples.npz")
    out = []
    for row in g["meta"]:
        k, near, far, radius, N, H, W, focal, f64 = row
        out.append(dict(k=int(k), near=float(near), far=float(far), radius=float(radius), N=int(N), H=int(H),
                        W=int(W), focal=float(focal), f64=bool(f64)))
    return g, out


def render_meta():
    g = load("render_grads.npz")
    out = []
    for row in g["meta"]:
        k, near, far, radius, N, H, W, n_codes, white = row
        out.append(dict(k=int(k), near=float(near), far=float(far), radius=float(radius), N=int(N), H=int(H),
                        W=int(W), n_codes=int(n_codes), white=bool(white)))
    return g, out


def render_case_inputs(c):
    """Re-create the inputs of render case `c` exactly as make_golden.py did."""
    k = c["k"]
    R = c["H"] * c["W"]
    return dict(
        c2w=syn.look_at_pose(200 + k, c["radius"]),
        focal=syn.SRN_FOCAL * c["W"] / 128.0,
        shape_codes=syn.make_codes(300 + k, c["n_codes"]),
        tex_codes=syn.make_codes(400 + k, c["n_codes"]),
        targets=syn.make_targets(500 + k, R),
        seed=2000 + k,
        R=R,
    )


def loss_seeds(rgb, depth, tgt, d_depth_coef=0.37):
    """d(loss)/d(rgb), d(loss)/d(depth) for loss = mean((rgb-tgt)^2) + coef*mean(depth)."""
    R = rgb.shape[0]
    d_rgb = (2.0 * (rgb - tgt) / (3.0 * R)).astype(np.float32)
    d_depth = np.full(R, d_depth_coef / R, np.float32)
    return d_rgb, d_depth


def reg_code_grad(codes, n_rows_total, coef=1e-4):
    """Gradient of coef*mean(norm(shape)+norm(tex)) wrt one code table (trainer.py:77-78).
    With per-ray codes every code row appears n_rows_total/n_codes times in the mean."""
    nrm = np.linalg.norm(codes.astype(np.float64), axis=-1, keepdims=True)
    return (coef * codes / nrm / codes.shape[0]).astype(np.float32)
