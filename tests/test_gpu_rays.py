"""get_rays / sample_from_rays on the GPU: bit-exact against the reference-generated golden
fixtures and against the oracle at the full 128x128 view size."""
import numpy as np
import pytest
import torch

from codenerf_b200 import synthetic as syn
from oracle import oracle as orc
from tests import golden_util as gu

G_RAYS, RAY_CASES = gu.rays_meta()


@pytest.mark.gpu
@pytest.mark.parametrize("c", RAY_CASES, ids=lambda c: f"c{c['k']}_{c['H']}x{c['W']}_N{c['N']}_{'f64' if c['f64'] else 'f32'}")
def test_rays_and_samples_bit_exact_vs_reference(c):
    import codenerf_b200 as cn
    k = c["k"]
    c2w = torch.from_numpy(syn.look_at_pose(100 + k, c["radius"]))
    focal = torch.tensor([c["focal"]], dtype=torch.float64) if c["f64"] else c["focal"]
    ro, vd = cn.get_rays(c["H"], c["W"], focal, c2w)
    assert ro.is_cuda and ro.shape == (c["H"] * c["W"], 3)
    ro_n, vd_n = ro.cpu().numpy(), vd.cpu().numpy()
    sub = np.arange(0, c["H"] * c["W"], 389)
    assert np.array_equal(gu.bits(ro_n[sub]), G_RAYS[f"c{k}_rays_o"])
    assert np.array_equal(gu.bits(vd_n[::13]), G_RAYS[f"c{k}_viewdirs_sub"])
    assert gu.checksum(vd_n) == G_RAYS[f"c{k}_viewdirs_sum"][0]
    torch.manual_seed(1000 + k)
    xyz, vdr, z = cn.sample_from_rays(ro, vd, c["near"], c["far"], c["N"])
    assert np.array_equal(gu.bits(z.cpu().numpy()), G_RAYS[f"c{k}_z"])
    xyz_n = xyz.cpu().numpy()
    assert np.array_equal(gu.bits(xyz_n[sub]), G_RAYS[f"c{k}_xyz_sub"])
    assert gu.checksum(xyz_n) == G_RAYS[f"c{k}_xyz_sum"][0]
    assert torch.equal(vdr[:, 0], vd) and torch.equal(vdr[:, -1], vd)
    zf = cn.sample_from_rays(ro[:1], vd[:1], c["near"], c["far"], c["N"], z_fixed=True)[2]
    assert np.array_equal(gu.bits(zf.cpu().numpy()), G_RAYS[f"c{k}_zfixed"])


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_rays_bit_exact_vs_oracle_random_focal(seed):
    import codenerf_b200 as cn
    H = W = 128
    focal = 100.0 + 37.123456789 * seed
    c2w = syn.look_at_pose(900 + seed, 1.3)
    for f64 in (True, False):
        ro_o, vd_o = orc.get_rays(H, W, focal, c2w, focal_is_f64=f64)
        farg = torch.tensor([focal], dtype=torch.float64) if f64 else focal
        ro, vd = cn.get_rays(H, W, farg, torch.from_numpy(c2w))
        assert np.array_equal(gu.bits(ro.cpu().numpy()), gu.bits(ro_o))
        assert np.array_equal(gu.bits(vd.cpu().numpy()), gu.bits(vd_o))
