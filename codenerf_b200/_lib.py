"""ctypes binding of libcodenerf_b200.so (the C ABI in include/codenerf_b200.h).

There is no CPU fallback and no alternative backend: if the shared library is missing, or no
B200-class CUDA device is present when an op is called, this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcodenerf_b200.so")
# CNB_LIB=trace selects the instrumented build (`python -m codenerf_b200.build --trace`); perf-debugging scripts only
if os.environ.get("CNB_LIB"):       # "trace", or an experiment variant built with `build.py --variant=<name> -D...`
    LIB_PATH = os.path.join(_HERE, "libcodenerf_b200_%s.so" % os.environ["CNB_LIB"])

PRECISION_BF16 = 0
PRECISION_FP32 = 1
_PRECISIONS = {"bf16": PRECISION_BF16, "fp32": PRECISION_FP32}


class NetConfig(ctypes.Structure):
    """cnb_net_config == CodeNeRF.__init__ kwargs (reference src/model.py:11-12)."""
    _fields_ = [("shape_blocks", ctypes.c_int32), ("texture_blocks", ctypes.c_int32), ("W", ctypes.c_int32),
                ("num_xyz_freq", ctypes.c_int32), ("num_dir_freq", ctypes.c_int32), ("latent_dim", ctypes.c_int32)]


class RayBatch(ctypes.Structure):
    """cnb_ray_batch (include/codenerf_b200.h)."""
    _fields_ = [
        ("n_rays", ctypes.c_int64),
        ("rays_per_segment", ctypes.c_int32),
        ("n_samples", ctypes.c_int32),
        ("rays_o", ctypes.c_void_p),
        ("viewdirs", ctypes.c_void_p),
        ("c2w", ctypes.c_void_p),
        ("pix_begin", ctypes.c_void_p),
        ("focal", ctypes.c_double),
        ("focal_is_f64", ctypes.c_int32),
        ("H", ctypes.c_int32),
        ("W", ctypes.c_int32),
        ("z_vals", ctypes.c_void_p),
        ("z_per_segment", ctypes.c_int32),
        ("segments_per_code", ctypes.c_int32),
        ("shape_codes", ctypes.c_void_p),
        ("texture_codes", ctypes.c_void_p),
        ("n_codes", ctypes.c_int32),
        ("white_bg", ctypes.c_int32),
    ]


EXPORTS = [
    "cnb_version", "cnb_strerror", "cnb_check_device", "cnb_param_count", "cnb_num_param_tensors",
    "cnb_param_layout", "cnb_get_rays", "cnb_sample_from_rays", "cnb_volume_rendering_forward",
    "cnb_volume_rendering_backward", "cnb_packed_weights_bytes", "cnb_pack_weights", "cnb_mlp_workspace_bytes",
    "cnb_mlp_forward", "cnb_mlp_backward", "cnb_render_workspace_bytes", "cnb_render_forward",
    "cnb_render_backward", "cnb_render_train_step", "cnb_launch_count", "cnb_debug_pipeline_timeouts",
    "cnb_profile_enable", "cnb_profile_read", "cnb_set_option", "cnb_clear_option", "cnb_get_option",
]

_lib = None


def load():
    """Load the shared library (no GPU needed to load; needed to call compute entry points)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m codenerf_b200.build` "
            "(nvcc, sm_100a). codenerf_b200 has no CPU or PyTorch fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, f32, f64, sz = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double,
                                  ctypes.c_size_t)
    cfgp, rayp = ctypes.POINTER(NetConfig), ctypes.POINTER(RayBatch)
    L.cnb_version.restype = i32
    L.cnb_strerror.restype = ctypes.c_char_p
    L.cnb_strerror.argtypes = [i32]
    L.cnb_check_device.restype = i32
    L.cnb_param_count.restype = i64
    L.cnb_param_count.argtypes = [cfgp]
    L.cnb_num_param_tensors.restype = i32
    L.cnb_num_param_tensors.argtypes = [cfgp]
    L.cnb_param_layout.restype = i32
    L.cnb_param_layout.argtypes = [cfgp, ctypes.POINTER(i64), ctypes.POINTER(ctypes.c_int32),
                                   ctypes.POINTER(ctypes.c_int32)]
    L.cnb_get_rays.restype = i32
    L.cnb_get_rays.argtypes = [i32, i32, f64, i32, vp, vp, vp, vp]
    L.cnb_sample_from_rays.restype = i32
    L.cnb_sample_from_rays.argtypes = [vp, vp, vp, i64, i32, vp, vp, vp]
    L.cnb_volume_rendering_forward.restype = i32
    L.cnb_volume_rendering_forward.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp, vp, vp]
    L.cnb_volume_rendering_backward.restype = i32
    L.cnb_volume_rendering_backward.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp, vp, vp, vp]
    L.cnb_packed_weights_bytes.restype = sz
    L.cnb_packed_weights_bytes.argtypes = [cfgp]
    L.cnb_pack_weights.restype = i32
    L.cnb_pack_weights.argtypes = [cfgp, ctypes.POINTER(vp), vp, vp]
    L.cnb_mlp_workspace_bytes.restype = sz
    L.cnb_mlp_workspace_bytes.argtypes = [cfgp, i64, i32, i32, i32]
    L.cnb_mlp_forward.restype = i32
    L.cnb_mlp_forward.argtypes = [cfgp, ctypes.POINTER(vp), vp, vp, vp, vp, vp, i32, i64, i64, i32, vp, vp, vp, sz, vp]
    L.cnb_mlp_backward.restype = i32
    L.cnb_mlp_backward.argtypes = [cfgp, ctypes.POINTER(vp), vp, vp, vp, vp, vp, i32, i64, i64, i32, vp, vp, vp, vp,
                                   vp, vp, sz, vp]
    L.cnb_render_workspace_bytes.restype = sz
    L.cnb_render_workspace_bytes.argtypes = [cfgp, rayp, i32, i32]
    L.cnb_render_forward.restype = i32
    L.cnb_render_forward.argtypes = [cfgp, ctypes.POINTER(vp), vp, rayp, i32, vp, vp, vp, vp, sz, vp]
    L.cnb_render_backward.restype = i32
    L.cnb_render_backward.argtypes = [cfgp, ctypes.POINTER(vp), vp, rayp, i32, vp, vp, vp, vp, vp, vp, sz, vp]
    L.cnb_render_train_step.restype = i32
    L.cnb_render_train_step.argtypes = [cfgp, ctypes.POINTER(vp), vp, rayp, i32, vp, f32, vp, vp, vp, vp, vp, vp, vp,
                                        vp, sz, vp]
    L.cnb_launch_count.restype = i64
    L.cnb_debug_pipeline_timeouts.restype = i32
    L.cnb_profile_enable.restype = i32
    L.cnb_profile_enable.argtypes = [i32]
    L.cnb_profile_read.restype = i32
    L.cnb_profile_read.argtypes = [i32, ctypes.POINTER(ctypes.c_float), i32]
    L.cnb_set_option.restype = i32
    L.cnb_set_option.argtypes = [ctypes.c_char_p, i64]
    L.cnb_clear_option.restype = i32
    L.cnb_clear_option.argtypes = [ctypes.c_char_p]
    L.cnb_get_option.restype = i64
    L.cnb_get_option.argtypes = [ctypes.c_char_p, i64]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise RuntimeError(f"codenerf_b200: {load().cnb_strerror(int(rc)).decode()} (status {rc})")


def set_option(name, value):
    """cnb_set_option: tuning / experiment switch (see include/codenerf_b200.h)."""
    check(load().cnb_set_option(name.encode(), int(value)))


def clear_option(name):
    check(load().cnb_clear_option(name.encode()))


def get_option(name, default=0):
    return int(load().cnb_get_option(name.encode(), int(default)))


def precision_id(p):
    if isinstance(p, int):
        return p
    try:
        return _PRECISIONS[p]
    except KeyError:
        raise ValueError(f"precision must be 'bf16' or 'fp32', got {p!r}") from None


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("codenerf_b200 needs a CUDA device (NVIDIA B200, sm_100); there is no CPU fallback")
