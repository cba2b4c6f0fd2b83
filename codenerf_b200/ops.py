"""Thin torch <-> C-ABI glue: device pointers, streams, workspaces, packed-weight cache.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); all arithmetic on
tensors happens in libcodenerf_b200.so.
"""
import ctypes

import torch

from . import _lib

_workspaces = {}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _f32c(t, device=None):
    """fp32, contiguous, on `device` (CUDA)."""
    if device is not None and t.device != device:
        t = t.to(device)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def workspace(device, nbytes):
    """Grow-only scratch (uint8, 256-B aligned by the caching allocator), one per (device, CUDA stream): calls on
    different streams never alias each other's scratch, calls on one stream are ordered by the stream.  A replaced
    (outgrown) tensor goes back to the caching allocator, which keeps it alive for the work already queued on its
    stream.  `release_workspaces()` drops them all (the training stash is ~1 MB per 128-sample tile)."""
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    w = _workspaces.get(key)
    if w is None or w.numel() < nbytes:
        w = None
        _workspaces.pop(key, None)
        w = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = w
    return w


def release_workspaces():
    """Free every cached scratch buffer (they are re-created on demand)."""
    _workspaces.clear()


_tuned = set()


def tune_for_device(device):
    """Once per device: size the backward sub-batches to the memory that is there.  Larger sub-batches mean fewer
    split-K flushes of the weight gradient (one 65,536-ray x 64-sample step: 16.3 ms at 8192 tiles, 15.65 ms in one
    32768-tile launch) and cost ~1 MB of workspace per tile; the library default (8192 tiles, 8 GB) is kept unless a
    third of the free memory covers more.  An explicit `cnb_set_option("sub_tiles")` / CNB_SUB_TILES wins."""
    key = (device.type, device.index)
    if key in _tuned:
        return
    _tuned.add(key)
    if _lib.get_option("sub_tiles", 0) != 0:
        return
    free, _total = torch.cuda.mem_get_info(device)
    for tiles in (32768, 16384):
        if tiles * (1100 << 10) <= free // 3:
            _lib.set_option("sub_tiles", tiles)
            return


def net_config(**kw):
    return _lib.NetConfig(**kw)


def param_pointer_table(params):
    arr = (ctypes.c_void_p * len(params))()
    for i, p in enumerate(params):
        if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
            raise RuntimeError("CodeNeRF parameters must be contiguous fp32 CUDA tensors")
        arr[i] = p.data_ptr()
    return arr


class PackedWeights:
    """bf16 tensor-core operand copies of the layer matrices (derived state).

    Rebuilt when a parameter's storage or version counter changes -- and, with `refresh=True`, unconditionally:
    fused optimisers (`torch.optim.AdamW(..., fused=True)`) and writes through `.data` update parameters WITHOUT
    bumping the version counter, so every caller whose parameters may be training passes `refresh=True` on the
    forward call (one ~15 us kernel).  Frozen models keep the cached copy; `invalidate()` drops it by hand."""

    def __init__(self):
        self.buf = None
        self.key = None

    def invalidate(self):
        self.key = None

    def get(self, cfg, params, refresh=False):
        key = tuple((p.data_ptr(), p._version) for p in params)
        if self.buf is None or self.key != key or refresh:
            L = _lib.load()
            n = L.cnb_packed_weights_bytes(ctypes.byref(cfg))
            if self.buf is None or self.buf.numel() < n or self.buf.device != params[0].device:
                self.buf = torch.empty(int(n), dtype=torch.uint8, device=params[0].device)
            _lib.check(L.cnb_pack_weights(ctypes.byref(cfg), param_pointer_table(params), _ptr(self.buf), _stream()))
            self.key = key
        return self.buf


# ---------------------------------------------------------------------------
def get_rays(H, W, focal, focal_is_f64, c2w):
    L = _lib.load()
    c2w = _f32c(c2w).reshape(4, 4)
    ro = torch.empty(H * W, 3, dtype=torch.float32, device=c2w.device)
    vd = torch.empty(H * W, 3, dtype=torch.float32, device=c2w.device)
    with torch.cuda.device(c2w.device):
        _lib.check(L.cnb_get_rays(int(H), int(W), float(focal), int(focal_is_f64), _ptr(c2w), _ptr(ro), _ptr(vd),
                                  _stream()))
    return ro, vd


def sample_from_rays(ro, vd, z_vals):
    L = _lib.load()
    R, N = ro.shape[0], z_vals.numel()
    xyz = torch.empty(R, N, 3, dtype=torch.float32, device=ro.device)
    vdr = torch.empty(R, N, 3, dtype=torch.float32, device=ro.device)
    with torch.cuda.device(ro.device):
        _lib.check(L.cnb_sample_from_rays(_ptr(ro), _ptr(vd), _ptr(z_vals), R, N, _ptr(xyz), _ptr(vdr), _stream()))
    return xyz, vdr


def volume_rendering_forward(sigmas, rgbs, z_vals, white_bg):
    L = _lib.load()
    N = z_vals.numel()
    B = sigmas.numel() // N
    rgb = torch.empty(B, 3, dtype=torch.float32, device=sigmas.device)
    depth = torch.empty(B, dtype=torch.float32, device=sigmas.device)
    acc = torch.empty(B, dtype=torch.float32, device=sigmas.device)
    if B == 0:
        return rgb, depth, acc
    with torch.cuda.device(sigmas.device):
        _lib.check(L.cnb_volume_rendering_forward(_ptr(sigmas), _ptr(rgbs), _ptr(z_vals), B, N, int(bool(white_bg)),
                                                  _ptr(rgb), _ptr(depth), _ptr(acc), _stream()))
    return rgb, depth, acc


def volume_rendering_backward(sigmas, rgbs, z_vals, white_bg, d_rgb, d_depth):
    L = _lib.load()
    N = z_vals.numel()
    B = sigmas.numel() // N
    ds = torch.empty(B, N, dtype=torch.float32, device=sigmas.device)
    dc = torch.empty(B, N, 3, dtype=torch.float32, device=sigmas.device)
    if B == 0:
        return ds, dc
    with torch.cuda.device(sigmas.device):
        _lib.check(L.cnb_volume_rendering_backward(_ptr(sigmas), _ptr(rgbs), _ptr(z_vals), B, N, int(bool(white_bg)),
                                                   _ptr(d_rgb), _ptr(d_depth), _ptr(ds), _ptr(dc), _stream()))
    return ds, dc


def mlp_forward(cfg, params, packed, xyz, viewdir, shape_codes, tex_codes, samples_per_code, precision):
    L = _lib.load()
    S = xyz.numel() // 3
    n_codes = shape_codes.shape[0]
    dev = xyz.device
    sig = torch.empty(S, dtype=torch.float32, device=dev)
    col = torch.empty(S, 3, dtype=torch.float32, device=dev)
    if S == 0:
        return sig, col
    with torch.cuda.device(dev):
        nws = L.cnb_mlp_workspace_bytes(ctypes.byref(cfg), S, n_codes, precision, 0)
        ws = workspace(dev, nws)
        _lib.check(L.cnb_mlp_forward(ctypes.byref(cfg), param_pointer_table(params), _ptr(packed), _ptr(xyz),
                                     _ptr(viewdir), _ptr(shape_codes), _ptr(tex_codes), n_codes,
                                     int(samples_per_code), S, precision, _ptr(sig), _ptr(col), _ptr(ws), ws.numel(),
                                     _stream()))
    return sig, col


def mlp_backward(cfg, params, packed, xyz, viewdir, shape_codes, tex_codes, samples_per_code, precision, d_sig,
                 d_col, want_param_grads):
    L = _lib.load()
    S = xyz.numel() // 3
    n_codes = shape_codes.shape[0]
    dev = xyz.device
    n_par = L.cnb_param_count(ctypes.byref(cfg))
    dP = torch.zeros(n_par, dtype=torch.float32, device=dev) if want_param_grads else None
    dsc = torch.empty_like(shape_codes)
    dtc = torch.empty_like(tex_codes)
    tune_for_device(dev)
    with torch.cuda.device(dev):
        nws = L.cnb_mlp_workspace_bytes(ctypes.byref(cfg), S, n_codes, precision, 1)
        ws = workspace(dev, nws)
        _lib.check(L.cnb_mlp_backward(ctypes.byref(cfg), param_pointer_table(params), _ptr(packed), _ptr(xyz),
                                      _ptr(viewdir), _ptr(shape_codes), _ptr(tex_codes), n_codes,
                                      int(samples_per_code), S, precision, _ptr(d_sig), _ptr(d_col), _ptr(dP),
                                      _ptr(dsc), _ptr(dtc), _ptr(ws), ws.numel(), _stream()))
    return dP, dsc, dtc


def split_flat_grads(cfg, flat, params):
    """Views of the flat state_dict-ordered gradient vector, one per parameter."""
    L = _lib.load()
    n = L.cnb_num_param_tensors(ctypes.byref(cfg))
    offs = (ctypes.c_int64 * n)()
    rows = (ctypes.c_int32 * n)()
    cols = (ctypes.c_int32 * n)()
    _lib.check(L.cnb_param_layout(ctypes.byref(cfg), offs, rows, cols))
    out = []
    for i, p in enumerate(params):
        out.append(flat[offs[i]:offs[i] + p.numel()].view_as(p))
    return out


class RayBatchArgs:
    """Keeps the tensors referenced by a cnb_ray_batch alive for the duration of a call."""

    def __init__(self, *, n_rays, rays_per_segment, n_samples, z_vals, shape_codes, tex_codes, white_bg=True,
                 rays_o=None, viewdirs=None, c2w=None, pix_begin=None, focal=0.0, focal_is_f64=True, H=0, W=0,
                 segments_per_code=1):
        dev = z_vals.device
        self.keep = []

        def prep(t, dtype=torch.float32):
            if t is None:
                return None
            t = t.to(device=dev, dtype=dtype).contiguous()
            self.keep.append(t)
            return t

        self.z_vals = prep(z_vals)
        self.shape_codes = prep(shape_codes.reshape(-1, shape_codes.shape[-1]))
        self.tex_codes = prep(tex_codes.reshape(-1, tex_codes.shape[-1]))
        self.rays_o = prep(rays_o)
        self.viewdirs = prep(viewdirs)
        self.c2w = prep(c2w)
        self.pix_begin = prep(pix_begin, torch.int32)
        n_seg = n_rays // rays_per_segment
        z_rows = self.z_vals.numel() // n_samples
        if z_rows not in (1, n_seg):
            raise ValueError("z_vals must have one row, or one row per segment")
        b = _lib.RayBatch()
        b.n_rays, b.rays_per_segment, b.n_samples = int(n_rays), int(rays_per_segment), int(n_samples)
        b.rays_o, b.viewdirs = (self.rays_o.data_ptr() if self.rays_o is not None else None,
                                self.viewdirs.data_ptr() if self.viewdirs is not None else None)
        b.c2w = self.c2w.data_ptr() if self.c2w is not None else None
        b.pix_begin = self.pix_begin.data_ptr() if self.pix_begin is not None else None
        b.focal, b.focal_is_f64, b.H, b.W = float(focal), int(bool(focal_is_f64)), int(H), int(W)
        b.z_vals = self.z_vals.data_ptr()
        b.z_per_segment = 1 if (z_rows == n_seg and n_seg > 1) else 0
        b.segments_per_code = int(segments_per_code)
        b.shape_codes, b.texture_codes = self.shape_codes.data_ptr(), self.tex_codes.data_ptr()
        b.n_codes = int(self.shape_codes.shape[0])
        b.white_bg = int(bool(white_bg))
        self.struct = b
        self.device = dev
        self.n_rays = int(n_rays)
        self.n_segments = n_seg


def render_forward(cfg, params, packed, rb, precision):
    L = _lib.load()
    dev = rb.device
    rgb = torch.empty(rb.n_rays, 3, dtype=torch.float32, device=dev)
    depth = torch.empty(rb.n_rays, dtype=torch.float32, device=dev)
    acc = torch.empty(rb.n_rays, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        nws = L.cnb_render_workspace_bytes(ctypes.byref(cfg), ctypes.byref(rb.struct), precision, 0)
        ws = workspace(dev, nws)
        _lib.check(L.cnb_render_forward(ctypes.byref(cfg), param_pointer_table(params), _ptr(packed),
                                        ctypes.byref(rb.struct), precision, _ptr(rgb), _ptr(depth), _ptr(acc),
                                        _ptr(ws), ws.numel(), _stream()))
    return rgb, depth, acc


def render_backward(cfg, params, packed, rb, precision, d_rgb, d_depth, want_param_grads):
    L = _lib.load()
    dev = rb.device
    n_par = L.cnb_param_count(ctypes.byref(cfg))
    dP = torch.zeros(n_par, dtype=torch.float32, device=dev) if want_param_grads else None
    dsc = torch.empty_like(rb.shape_codes)
    dtc = torch.empty_like(rb.tex_codes)
    tune_for_device(dev)
    with torch.cuda.device(dev):
        nws = L.cnb_render_workspace_bytes(ctypes.byref(cfg), ctypes.byref(rb.struct), precision, 1)
        ws = workspace(dev, nws)
        _lib.check(L.cnb_render_backward(ctypes.byref(cfg), param_pointer_table(params), _ptr(packed),
                                         ctypes.byref(rb.struct), precision, _ptr(d_rgb), _ptr(d_depth), _ptr(dP),
                                         _ptr(dsc), _ptr(dtc), _ptr(ws), ws.numel(), _stream()))
    return dP, dsc, dtc


def render_train_step(cfg, params, packed, rb, precision, target, loss_scale, d_params, want_outputs=True):
    """Forward + per-segment L2 loss + backward.  d_params (flat fp32 or None) is accumulated into."""
    L = _lib.load()
    dev = rb.device
    rgb = depth = acc = None
    if want_outputs:
        rgb = torch.empty(rb.n_rays, 3, dtype=torch.float32, device=dev)
        depth = torch.empty(rb.n_rays, dtype=torch.float32, device=dev)
        acc = torch.empty(rb.n_rays, dtype=torch.float32, device=dev)
    sq = torch.empty(rb.n_segments, dtype=torch.float32, device=dev)
    dsc = torch.empty_like(rb.shape_codes)
    dtc = torch.empty_like(rb.tex_codes)
    tune_for_device(dev)
    with torch.cuda.device(dev):
        nws = L.cnb_render_workspace_bytes(ctypes.byref(cfg), ctypes.byref(rb.struct), precision, 1)
        ws = workspace(dev, nws)
        _lib.check(L.cnb_render_train_step(ctypes.byref(cfg), param_pointer_table(params), _ptr(packed),
                                           ctypes.byref(rb.struct), precision, _ptr(target), float(loss_scale),
                                           _ptr(rgb), _ptr(depth), _ptr(acc), _ptr(sq), _ptr(d_params), _ptr(dsc),
                                           _ptr(dtc), _ptr(ws), ws.numel(), _stream()))
    return rgb, depth, acc, sq, dsc, dtc
