"""ctypes binding of the CPU oracle (oracle/codenerf_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(codenerf_b200/) never imports this module.

All arrays are numpy float32, C-contiguous.  Function names mirror the reference
(yuliangguo/code-nerf): src/utils.py get_rays / sample_from_rays /
volume_rendering, src/model.py PE / CodeNeRF.forward.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcodenerf_oracle.so")


def build(force=False):
    """Compile the C restatement (gcc, a second or two)."""
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "codenerf_oracle.c"))):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libcodenerf_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


class OrcConfig(ctypes.Structure):
    _fields_ = [("shape_blocks", ctypes.c_int), ("texture_blocks", ctypes.c_int), ("W", ctypes.c_int),
                ("num_xyz_freq", ctypes.c_int), ("num_dir_freq", ctypes.c_int), ("latent_dim", ctypes.c_int)]


SRN_CONFIG = dict(shape_blocks=3, texture_blocks=1, W=256, num_xyz_freq=10, num_dir_freq=4, latent_dim=256)

_lib = None
_fp = ctypes.POINTER(ctypes.c_float)


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.orc_param_count.restype = ctypes.c_int64
        L.orc_param_count.argtypes = [ctypes.POINTER(OrcConfig)]
        L.orc_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_fp)


def _cfg(cfg):
    c = dict(SRN_CONFIG)
    if cfg:
        c.update(cfg)
    return OrcConfig(**c)


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    """OpenMP threads of the port (bench.py: torchrun pins OMP_NUM_THREADS=1; the CPU baseline may use every core)."""
    lib().orc_set_num_threads(ctypes.c_int(int(n)))


def param_count(cfg=None):
    c = _cfg(cfg)
    return int(lib().orc_param_count(ctypes.byref(c)))


def param_shapes(cfg=None):
    """[(state_dict key, shape)] in CodeNeRF.state_dict() order (src/model.py:20-34)."""
    c = dict(SRN_CONFIG)
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


def split_params(flat, cfg=None):
    """flat fp32 vector -> {state_dict key: array view}."""
    out, o = {}, 0
    for k, shp in param_shapes(cfg):
        n = int(np.prod(shp))
        out[k] = flat[o:o + n].reshape(shp)
        o += n
    assert o == flat.size
    return out


def torch_rand(seed, n):
    """torch.manual_seed(seed); torch.rand(n) (CPU mt19937) -- src/utils.py:29."""
    out = np.empty(n, np.float32)
    lib().orc_torch_rand(ctypes.c_uint32(seed), ctypes.c_int(n), out.ctypes.data_as(_fp))
    return out


def get_rays(H, W, focal, c2w, focal_is_f64=True):
    """src/utils.py:10-19 -> (rays_o [HW,3], viewdirs [HW,3])."""
    c2w, pc = _f(np.asarray(c2w).reshape(4, 4))
    ro = np.empty((H * W, 3), np.float32)
    vd = np.empty((H * W, 3), np.float32)
    lib().orc_get_rays(ctypes.c_int(H), ctypes.c_int(W), ctypes.c_double(float(focal)),
                       ctypes.c_int(1 if focal_is_f64 else 0), pc, ro.ctypes.data_as(_fp), vd.ctypes.data_as(_fp))
    return ro, vd


def z_vals(near, far, N, rnd=None, z_fixed=False):
    """z_vals of src/utils.py:24-29; rnd = the torch.rand(N) draw."""
    z = np.empty(N, np.float32)
    if z_fixed:
        lib().orc_z_vals(ctypes.c_double(near), ctypes.c_double(far), ctypes.c_int(N), None, ctypes.c_int(1),
                         z.ctypes.data_as(_fp))
    else:
        rnd, pr = _f(rnd)
        assert rnd.size == N
        lib().orc_z_vals(ctypes.c_double(near), ctypes.c_double(far), ctypes.c_int(N), pr, ctypes.c_int(0),
                         z.ctypes.data_as(_fp))
    return z


def sample_from_rays(ro, vd, z):
    """xyz / repeated viewdir of src/utils.py:30-31."""
    ro, pro = _f(ro)
    vd, pvd = _f(vd)
    z, pz = _f(z)
    R, N = ro.shape[0], z.size
    xyz = np.empty((R, N, 3), np.float32)
    vdr = np.empty((R, N, 3), np.float32)
    lib().orc_sample_from_rays(pro, pvd, pz, ctypes.c_int(R), ctypes.c_int(N), xyz.ctypes.data_as(_fp),
                               vdr.ctypes.data_as(_fp))
    return xyz, vdr


def pe(x, degree):
    """src/model.py:4-7."""
    x, px = _f(x)
    n = x.size // 3
    out = np.empty(x.shape[:-1] + (3 + 6 * degree,), np.float32)
    lib().orc_pe(px, ctypes.c_int(n), ctypes.c_int(degree), out.ctypes.data_as(_fp))
    return out


def _codes(shape_codes, tex_codes):
    sc, psc = _f(np.atleast_2d(shape_codes))
    tc, ptc = _f(np.atleast_2d(tex_codes))
    assert sc.shape == tc.shape
    return sc, psc, tc, ptc


def mlp_forward(params_flat, xyz, viewdir, shape_codes, tex_codes, samples_per_code=0, cfg=None):
    """CodeNeRF.forward (src/model.py:36-53).  xyz/viewdir [...,3] -> sigmas [...,1], rgbs [...,3]."""
    c = _cfg(cfg)
    P, pP = _f(params_flat)
    xyz, px = _f(xyz)
    vd, pv = _f(viewdir)
    sc, psc, tc, ptc = _codes(shape_codes, tex_codes)
    S = xyz.size // 3
    sig = np.empty(xyz.shape[:-1] + (1,), np.float32)
    rgb = np.empty(xyz.shape[:-1] + (3,), np.float32)
    rc = lib().orc_mlp_forward(ctypes.byref(c), pP, px, pv, psc, ptc, ctypes.c_int(sc.shape[0]),
                               ctypes.c_int64(samples_per_code), ctypes.c_int64(S),
                               sig.ctypes.data_as(_fp), rgb.ctypes.data_as(_fp))
    assert rc == 0
    return sig, rgb


def volume_rendering(sigmas, rgbs, z, white_bg=True):
    """src/utils.py:34-47 -> (rgb [B,3], depth [B], acc [B])."""
    sg, ps = _f(sigmas)
    cl, pc = _f(rgbs)
    z, pz = _f(z)
    N = z.size
    B = sg.size // N
    rgb = np.empty((B, 3), np.float32)
    depth = np.empty(B, np.float32)
    acc = np.empty(B, np.float32)
    lib().orc_volume_rendering(ps, pc, pz, ctypes.c_int64(B), ctypes.c_int(N), ctypes.c_int(1 if white_bg else 0),
                               rgb.ctypes.data_as(_fp), depth.ctypes.data_as(_fp), acc.ctypes.data_as(_fp))
    return rgb, depth, acc


def volume_rendering_backward(sigmas, rgbs, z, d_rgb, d_depth=None, white_bg=True):
    sg, ps = _f(sigmas)
    cl, pc = _f(rgbs)
    z, pz = _f(z)
    N = z.size
    B = sg.size // N
    g, pg = _f(d_rgb)
    pd = None
    if d_depth is not None:
        gd, pd = _f(d_depth)
    ds = np.empty((B, N), np.float32)
    dc = np.empty((B, N, 3), np.float32)
    lib().orc_volume_rendering_backward(ps, pc, pz, ctypes.c_int64(B), ctypes.c_int(N),
                                        ctypes.c_int(1 if white_bg else 0), pg, pd,
                                        ds.ctypes.data_as(_fp), dc.ctypes.data_as(_fp))
    return ds, dc


def mlp_backward(params_flat, xyz, viewdir, shape_codes, tex_codes, d_sigmas, d_rgbs, samples_per_code=0,
                 cfg=None, want_param_grads=True):
    """Autograd of CodeNeRF.forward -> (dP flat, d_shape_codes, d_tex_codes)."""
    c = _cfg(cfg)
    P, pP = _f(params_flat)
    xyz, px = _f(xyz)
    vd, pv = _f(viewdir)
    sc, psc, tc, ptc = _codes(shape_codes, tex_codes)
    ds, pds = _f(d_sigmas)
    dc, pdc = _f(d_rgbs)
    S = xyz.size // 3
    dP = np.zeros(P.size, np.float32) if want_param_grads else None
    dsc = np.empty_like(sc)
    dtc = np.empty_like(tc)
    rc = lib().orc_mlp_backward(ctypes.byref(c), pP, px, pv, psc, ptc, ctypes.c_int(sc.shape[0]),
                                ctypes.c_int64(samples_per_code), ctypes.c_int64(S), pds, pdc,
                                dP.ctypes.data_as(_fp) if dP is not None else None,
                                dsc.ctypes.data_as(_fp), dtc.ctypes.data_as(_fp))
    assert rc == 0
    return dP, dsc, dtc


# ---------------------------------------------------------------------------
# Whole-path helpers (the 4-call idiom of src/trainer.py:65-82)

def render(params_flat, H, W, focal, c2w, z, shape_codes, tex_codes, white_bg=True, cfg=None,
           ray_begin=0, ray_count=None, focal_is_f64=True):
    """get_rays -> sample_from_rays -> CodeNeRF -> volume_rendering for rays
    [ray_begin, ray_begin+ray_count) of one view.  Returns dict with rgb/depth/acc
    and the intermediates needed by render_backward."""
    ro, vd = get_rays(H, W, focal, c2w, focal_is_f64)
    if ray_count is None:
        ray_count = H * W - ray_begin
    ro, vd = ro[ray_begin:ray_begin + ray_count], vd[ray_begin:ray_begin + ray_count]
    xyz, vdr = sample_from_rays(ro, vd, z)
    sig, col = mlp_forward(params_flat, xyz, vdr, shape_codes, tex_codes, 0, cfg)
    rgb, depth, acc = volume_rendering(sig, col, z, white_bg)
    return dict(rgb=rgb, depth=depth, acc=acc, xyz=xyz, vd=vdr, sigmas=sig, rgbs=col)


def render_backward(params_flat, fwd, z, shape_codes, tex_codes, d_rgb, d_depth=None, white_bg=True, cfg=None,
                    want_param_grads=True):
    ds, dc = volume_rendering_backward(fwd["sigmas"], fwd["rgbs"], z, d_rgb, d_depth, white_bg)
    return mlp_backward(params_flat, fwd["xyz"], fwd["vd"], shape_codes, tex_codes, ds, dc, 0, cfg,
                        want_param_grads)
