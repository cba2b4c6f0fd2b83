"""Checkpoint wire formats: models.pth / codes.pth round trips, key sets, and -- when the reference
checkout is present (authoring container only) -- a cross-load with the reference's own nn.Module."""
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from codenerf_b200 import checkpoint as ck  # noqa: E402
from codenerf_b200 import synthetic as syn  # noqa: E402
from codenerf_b200.model import CodeNeRF  # noqa: E402

REF_KEYS_SHAPES = {       # reference src/model.py:11-34
    "encoding_xyz.0.weight": (256, 63), "encoding_xyz.0.bias": (256,),
    "shape_latent_layer_1.0.weight": (256, 256), "shape_layer_1.0.weight": (256, 256),
    "encoding_shape.weight": (256, 256), "sigma.0.weight": (1, 256),
    "encoding_viewdir.0.weight": (256, 283), "texture_latent_layer_1.0.weight": (256, 256),
    "texture_layer_1.0.weight": (256, 256), "rgb.0.weight": (128, 256), "rgb.2.weight": (3, 128), "rgb.2.bias": (3,),
}


def _model(seed=0):
    flat, views = syn.make_params(seed)
    m = CodeNeRF(**syn.SRN_NET, precision="fp32")
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in views.items()})
    return m


def test_models_pth_round_trip(tmp_path):
    m = _model(1)
    sc = torch.nn.Embedding(7, 256); tc = torch.nn.Embedding(7, 256)
    d = ck.save_models(str(tmp_path), m, sc, tc, niter=123, nepoch=4, iteration=100)
    assert tuple(d.keys()) == ck.MODEL_KEYS
    assert os.path.exists(tmp_path / "models.pth") and os.path.exists(tmp_path / "100.pth")
    sd = d["model_params"]
    assert len(sd) == 28 and sum(v.numel() for v in sd.values()) == 714756
    for k, shp in REF_KEYS_SHAPES.items():
        assert tuple(sd[k].shape) == shp, k
    m2 = CodeNeRF(**syn.SRN_NET, precision="fp32")
    saved, mean_s, mean_t = ck.load_models(str(tmp_path / "models.pth"), m2)
    assert saved["niter"] == 123 and saved["nepoch"] == 4
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    assert mean_s.shape == (1, 256) and torch.allclose(mean_s[0], sc.weight.detach().mean(0))
    assert torch.allclose(mean_t[0], tc.weight.detach().mean(0))


def test_load_rejects_foreign_files(tmp_path):
    torch.save({"model_params": {}}, tmp_path / "x.pth")
    with pytest.raises(KeyError):
        ck.load_models(str(tmp_path / "x.pth"))
    m = _model(0)
    bad = ck.models_dict(m, torch.zeros(2, 256), torch.zeros(2, 256), 0, 0)
    bad["model_params"] = {k: v for k, v in bad["model_params"].items() if not k.startswith("sigma")}
    torch.save(bad, tmp_path / "y.pth")
    with pytest.raises(RuntimeError):                     # strict load: the key set is part of the boundary
        ck.load_models(str(tmp_path / "y.pth"), CodeNeRF(**syn.SRN_NET, precision="fp32"))


def test_codes_pth_round_trip(tmp_path):
    ids = ["a1", "b2", "c3"]
    s, t = torch.randn(3, 256), torch.randn(3, 256)
    ck.save_codes(str(tmp_path), ids, 2, s, t, {0: [30.1], 1: [29.5]}, {0: [0.9]})
    d = ck.load_codes(str(tmp_path / "codes.pth"))
    assert tuple(d.keys()) == ck.CODES_KEYS and d["ids"] == ids and d["num_obj"] == 2
    assert torch.equal(d["optimized_shapecodes"], s) and torch.equal(d["optimized_texturecodes"], t)
    assert d["psnr_eval"][1] == [29.5]


@pytest.mark.skipif(not os.path.exists("/root/reference/src/model.py"), reason="reference checkout not present")
def test_cross_load_with_the_reference_module(tmp_path):
    """A models.pth written here loads into the reference's CodeNeRF and vice versa (strict), and both modules then
    compute the same sigmas / rgbs (fp32 mode is not needed for this: the comparison runs the REFERENCE forward on
    both parameter sets on the CPU)."""
    sys.modules.setdefault("imageio", types.ModuleType("imageio"))
    sys.path.insert(0, "/root/reference/src")
    try:
        import importlib
        ref_model = importlib.import_module("model")
    finally:
        sys.path.remove("/root/reference/src")
    ref = ref_model.CodeNeRF(**{k: v for k, v in syn.SRN_NET.items()})
    mine = _model(3)
    ck.save_models(str(tmp_path / "a"), mine, torch.zeros(2, 256), torch.zeros(2, 256), 1, 1)
    saved = torch.load(tmp_path / "a" / "models.pth", map_location="cpu", weights_only=False)
    ref.load_state_dict(saved["model_params"])                      # ours -> reference, strict
    ck.save_models(str(tmp_path / "b"), ref, torch.zeros(2, 256), torch.zeros(2, 256), 1, 1)
    back = CodeNeRF(**syn.SRN_NET, precision="fp32")
    ck.load_models(str(tmp_path / "b" / "models.pth"), back)         # reference -> ours, strict
    for (k1, v1), (k2, v2) in zip(mine.state_dict().items(), back.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
