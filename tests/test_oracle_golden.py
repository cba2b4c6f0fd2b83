"""The CPU oracle (oracle/codenerf_oracle.c) against fixtures produced by the unmodified
reference (tests/golden/make_golden.py).  Rays / samples: bit-exact.  Floating-point network
and compositing: 2e-5 absolute (fp32 summation order differs from ATen/MKL only)."""
import numpy as np
import pytest

from codenerf_b200 import synthetic as syn
from oracle import oracle as orc
from tests import golden_util as gu

G_RAYS, RAY_CASES = gu.rays_meta()
G_REN, REN_CASES = gu.render_meta()


@pytest.mark.parametrize("c", RAY_CASES, ids=lambda c: f"c{c['k']}_{c['H']}x{c['W']}_N{c['N']}_{'f64' if c['f64'] else 'f32'}")
def test_rays_and_samples_bit_exact(c):
    k = c["k"]
    c2w = syn.look_at_pose(100 + k, c["radius"])
    ro, vd = orc.get_rays(c["H"], c["W"], c["focal"], c2w, focal_is_f64=c["f64"])
    sub = np.arange(0, c["H"] * c["W"], 389)
    assert np.array_equal(gu.bits(ro[sub]), G_RAYS[f"c{k}_rays_o"])
    assert np.array_equal(gu.bits(vd[::13]), G_RAYS[f"c{k}_viewdirs_sub"])
    assert gu.checksum(vd) == G_RAYS[f"c{k}_viewdirs_sum"][0]
    if f"c{k}_viewdirs" in G_RAYS:
        assert np.array_equal(gu.bits(vd), G_RAYS[f"c{k}_viewdirs"])
    rnd = orc.torch_rand(1000 + k, c["N"])
    z = orc.z_vals(c["near"], c["far"], c["N"], rnd)
    assert np.array_equal(gu.bits(z), G_RAYS[f"c{k}_z"])
    zf = orc.z_vals(c["near"], c["far"], c["N"], z_fixed=True)
    assert np.array_equal(gu.bits(zf), G_RAYS[f"c{k}_zfixed"])
    xyz, vdr = orc.sample_from_rays(ro, vd, z)
    assert np.array_equal(gu.bits(xyz[sub]), G_RAYS[f"c{k}_xyz_sub"])
    assert gu.checksum(xyz) == G_RAYS[f"c{k}_xyz_sum"][0]
    assert np.array_equal(vdr[:, 0], vd) and np.array_equal(vdr[:, -1], vd)


def test_torch_rand_is_mt19937():
    import torch
    for seed, n in ((0, 64), (1234, 96), (2**31 + 5, 700)):
        torch.manual_seed(seed)
        assert np.array_equal(torch.rand(n).numpy(), orc.torch_rand(seed % 2**32, n))


def _forward(c):
    k = c["k"]
    inp = gu.render_case_inputs(c)
    flat, _ = syn.make_params(0)
    z = G_REN[f"r{k}_z"].view(np.float32)
    z2 = orc.z_vals(c["near"], c["far"], c["N"], orc.torch_rand(inp["seed"], c["N"]))
    assert np.array_equal(z, z2)
    ro, vd = orc.get_rays(c["H"], c["W"], inp["focal"], inp["c2w"], True)
    xyz, vdr = orc.sample_from_rays(ro, vd, z)
    spc = 0 if c["n_codes"] == 1 else (inp["R"] // c["n_codes"]) * c["N"]
    sig, col = orc.mlp_forward(flat, xyz, vdr, inp["shape_codes"], inp["tex_codes"], spc)
    return inp, flat, z, xyz, vdr, sig, col, spc


@pytest.mark.parametrize("c", REN_CASES, ids=lambda c: f"r{c['k']}")
def test_forward_matches_reference(c):
    k = c["k"]
    inp, flat, z, xyz, vdr, sig, col, spc = _forward(c)
    np.testing.assert_allclose(orc.pe(xyz[:3], 10), G_REN[f"r{k}_pe_xyz"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(orc.pe(vdr[:3], 4), G_REN[f"r{k}_pe_dir"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(sig.reshape(inp["R"], c["N"]), G_REN[f"r{k}_sigmas"], atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose(col, G_REN[f"r{k}_rgbs"], atol=2e-5, rtol=1e-5)
    rgb, depth, acc = orc.volume_rendering(sig, col, z, c["white"])
    np.testing.assert_allclose(rgb, G_REN[f"r{k}_rgb"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(depth, G_REN[f"r{k}_depth"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(acc, G_REN[f"r{k}_acc"], atol=2e-5, rtol=0)


@pytest.mark.parametrize("c", REN_CASES, ids=lambda c: f"r{c['k']}")
def test_backward_matches_reference_autograd(c):
    k = c["k"]
    inp, flat, z, xyz, vdr, sig, col, spc = _forward(c)
    rgb, depth, acc = orc.volume_rendering(sig, col, z, c["white"])
    d_rgb, d_depth = gu.loss_seeds(rgb, depth, inp["targets"])
    ds, dc = orc.volume_rendering_backward(sig, col, z, d_rgb, d_depth, c["white"])
    dP, dsc, dtc = orc.mlp_backward(flat, xyz, vdr, inp["shape_codes"], inp["tex_codes"], ds, dc, spc)
    dsc = dsc + gu.reg_code_grad(inp["shape_codes"], inp["R"])
    dtc = dtc + gu.reg_code_grad(inp["tex_codes"], inp["R"])
    scale = max(np.abs(G_REN[f"r{k}_d_shape"]).max(), 1e-12)
    np.testing.assert_allclose(dsc, G_REN[f"r{k}_d_shape"], atol=2e-4 * scale, rtol=1e-4)
    np.testing.assert_allclose(dtc, G_REN[f"r{k}_d_tex"], atol=2e-4 * scale, rtol=1e-4)
    grads = orc.split_params(dP)
    for t, (key, shp) in enumerate(orc.param_shapes()):
        g = grads[key]
        stat = G_REN[f"r{k}_g/stat/{key}"]
        g64 = g.astype(np.float64).ravel()
        mine = np.array([g64.sum(), np.abs(g64).sum(), float(g64 @ gu.weight_probe(t, g64.size))])
        tol = 2e-4 * stat[1] + 1e-9
        assert np.all(np.abs(mine - stat) <= tol), (key, mine, stat)
        if g.ndim == 1:
            ref = G_REN[f"r{k}_g/full/{key}"]
            np.testing.assert_allclose(g, ref, atol=2e-4 * max(np.abs(ref).max(), 1e-12), rtol=1e-4, err_msg=key)
        else:
            ref = G_REN[f"r{k}_g/head/{key}"]
            np.testing.assert_allclose(g[:4], ref, atol=2e-4 * max(np.abs(ref).max(), 1e-12), rtol=1e-4, err_msg=key)
