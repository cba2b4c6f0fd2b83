"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares.
No compute entry point is called here (no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for f in os.listdir(os.path.join(ROOT, "include")):
        if f.endswith(".h"):
            src = open(os.path.join(ROOT, "include", f)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(cnb_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_builds_and_exports_all_declared_symbols():
    from codenerf_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ but not exported"
    assert sorted(_lib.EXPORTS) == syms


def test_host_only_entry_points():
    from codenerf_b200 import _lib
    L = _lib.load()
    cfg = _lib.NetConfig(shape_blocks=3, texture_blocks=1, W=256, num_xyz_freq=10, num_dir_freq=4, latent_dim=256)
    assert L.cnb_param_count(ctypes.byref(cfg)) == 714756
    n = L.cnb_num_param_tensors(ctypes.byref(cfg))
    assert n == 28
    offs, rows, cols = (ctypes.c_int64 * n)(), (ctypes.c_int32 * n)(), (ctypes.c_int32 * n)()
    assert L.cnb_param_layout(ctypes.byref(cfg), offs, rows, cols) == 0
    assert (rows[0], cols[0]) == (256, 63) and (rows[18], cols[18]) == (256, 283)
    bad = _lib.NetConfig(shape_blocks=0, texture_blocks=1, W=256, num_xyz_freq=10, num_dir_freq=4, latent_dim=256)
    assert L.cnb_param_count(ctypes.byref(bad)) < 0
    assert b"invalid" in L.cnb_strerror(-1)
    assert L.cnb_packed_weights_bytes(ctypes.byref(cfg)) > 0


def test_options_registry():
    """cnb_set_option / cnb_get_option / cnb_clear_option: explicit value > CNB_<NAME> environment variable > default."""
    from codenerf_b200 import _lib
    L = _lib.load()
    old = os.environ.pop("CNB_SUB_TILES", None)
    try:
        assert _lib.get_option("sub_tiles", 8192) == 8192
        os.environ["CNB_SUB_TILES"] = "4096"
        assert _lib.get_option("sub_tiles", 8192) == 4096
        _lib.set_option("sub_tiles", 16384)
        assert _lib.get_option("sub_tiles", 8192) == 16384
        _lib.clear_option("sub_tiles")
        assert _lib.get_option("sub_tiles", 8192) == 4096
        assert L.cnb_set_option(b"no_such_option", 1) == -1
    finally:
        os.environ.pop("CNB_SUB_TILES", None)
        if old is not None:
            os.environ["CNB_SUB_TILES"] = old


def test_state_dict_matches_reference_layout():
    """Keys / shapes of reference src/model.py:20-34."""
    import codenerf_b200 as cn
    from codenerf_b200 import synthetic as syn
    m = cn.CodeNeRF(**syn.SRN_NET)
    sd = m.state_dict()
    assert [(k, tuple(v.shape)) for k, v in sd.items()] == syn.param_shapes()
    assert sum(v.numel() for v in sd.values()) == 714756


def test_no_cpu_fallback():
    import torch
    import codenerf_b200 as cn
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        cn.get_rays(4, 4, 10.0, torch.eye(4))
    m = cn.CodeNeRF(shape_blocks=3)
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 4, 3), torch.zeros(2, 4, 3), torch.zeros(1, 256), torch.zeros(1, 256))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "codenerf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "synthetic.py", f
