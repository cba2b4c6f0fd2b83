"""fp32 mode (CUDA cores) against the oracle and the reference-generated fixtures.
Tolerance 1e-4 absolute on rgb/depth/acc (BASELINE.json north_star, fp32-accumulate mode)."""
import numpy as np
import pytest
import torch

from codenerf_b200 import synthetic as syn
from oracle import oracle as orc
from tests import golden_util as gu
from tests import gpu_util as U

G_REN, REN_CASES = gu.render_meta()
ATOL = 1e-4


def _case_tensors(c):
    inp = gu.render_case_inputs(c)
    k = c["k"]
    z = G_REN[f"r{k}_z"].view(np.float32)
    ro, vd = orc.get_rays(c["H"], c["W"], inp["focal"], inp["c2w"], True)
    xyz, vdr = orc.sample_from_rays(ro, vd, z)
    return inp, z, xyz, vdr


@pytest.mark.gpu
@pytest.mark.parametrize("c", REN_CASES, ids=lambda c: f"r{c['k']}")
def test_unfused_api_matches_reference_fixture(c):
    """model(xyz, viewdir, codes) -> volume_rendering -> loss.backward(), the reference's own idiom."""
    import codenerf_b200 as cn
    k = c["k"]
    inp, z, xyz, vdr = _case_tensors(c)
    model, flat = U.make_model("fp32")
    R, N = inp["R"], c["N"]
    sc = torch.from_numpy(inp["shape_codes"]).cuda().requires_grad_()
    tc = torch.from_numpy(inp["tex_codes"]).cuda().requires_grad_()
    if c["n_codes"] == 1:
        sc_in, tc_in = sc, tc
    else:
        per = R // c["n_codes"]
        sc_in = sc.repeat_interleave(per, 0).unsqueeze(1)
        tc_in = tc.repeat_interleave(per, 0).unsqueeze(1)
    sig, col = model(torch.from_numpy(xyz).cuda(), torch.from_numpy(vdr).cuda(), sc_in, tc_in)
    assert sig.shape == (R, N, 1) and col.shape == (R, N, 3)
    np.testing.assert_allclose(sig.detach().cpu().numpy().reshape(R, N), G_REN[f"r{k}_sigmas"], atol=ATOL, rtol=1e-4)
    np.testing.assert_allclose(col.detach().cpu().numpy(), G_REN[f"r{k}_rgbs"], atol=ATOL, rtol=1e-4)
    rgb, depth, acc = cn.volume_rendering_with_acc(sig, col, torch.from_numpy(z).cuda(), white_bg=c["white"])
    np.testing.assert_allclose(rgb.detach().cpu().numpy(), G_REN[f"r{k}_rgb"], atol=ATOL, rtol=0)
    np.testing.assert_allclose(depth.detach().cpu().numpy(), G_REN[f"r{k}_depth"], atol=ATOL, rtol=0)
    np.testing.assert_allclose(acc.detach().cpu().numpy(), G_REN[f"r{k}_acc"], atol=ATOL, rtol=0)
    tgt = torch.from_numpy(inp["targets"]).cuda()
    loss_l2 = torch.mean((rgb - tgt) ** 2)
    reg = torch.norm(sc_in, dim=-1) + torch.norm(tc_in, dim=-1)
    loss = loss_l2 + 1e-4 * torch.mean(reg) + 0.37 * depth.mean()
    loss.backward()
    assert abs(loss.item() - G_REN[f"r{k}_loss"][1]) < 1e-5
    ref_s, ref_t = G_REN[f"r{k}_d_shape"], G_REN[f"r{k}_d_tex"]
    np.testing.assert_allclose(sc.grad.cpu().numpy(), ref_s, atol=5e-4 * np.abs(ref_s).max(), rtol=1e-3)
    np.testing.assert_allclose(tc.grad.cpu().numpy(), ref_t, atol=5e-4 * np.abs(ref_t).max(), rtol=1e-3)
    for t, (key, p) in enumerate(model.named_parameters()):
        g = p.grad.cpu().numpy()
        stat = G_REN[f"r{k}_g/stat/{key}"]
        g64 = g.astype(np.float64).ravel()
        mine = np.array([g64.sum(), np.abs(g64).sum(), float(g64 @ gu.weight_probe(t, g64.size))])
        assert np.all(np.abs(mine - stat) <= 5e-4 * stat[1] + 1e-9), (key, mine, stat)
        ref = G_REN[f"r{k}_g/full/{key}"] if g.ndim == 1 else G_REN[f"r{k}_g/head/{key}"]
        got = g if g.ndim == 1 else g[:4]
        np.testing.assert_allclose(got, ref, atol=5e-4 * max(np.abs(ref).max(), 1e-12), rtol=1e-3, err_msg=key)


@pytest.mark.gpu
@pytest.mark.parametrize("N,B", [(64, 300), (96, 77), (40, 33), (1, 5), (257, 9)])
def test_volume_rendering_vs_oracle(N, B):
    import codenerf_b200 as cn
    rng = np.random.default_rng(N * 1000 + B)
    sig = np.abs(rng.normal(size=(B, N, 1))).astype(np.float32) * 20
    col = rng.normal(size=(B, N, 3)).astype(np.float32)
    z = np.sort(rng.uniform(0.8, 1.8, N)).astype(np.float32)
    for white in (True, False):
        r0, d0, a0 = orc.volume_rendering(sig, col, z, white)
        s_t = torch.from_numpy(sig).cuda().requires_grad_()
        c_t = torch.from_numpy(col).cuda().requires_grad_()
        rgb, depth, acc = cn.volume_rendering_with_acc(s_t, c_t, torch.from_numpy(z).cuda(), white_bg=white)
        np.testing.assert_allclose(rgb.detach().cpu().numpy(), r0, atol=2e-5)
        np.testing.assert_allclose(depth.detach().cpu().numpy(), d0, atol=2e-5)
        np.testing.assert_allclose(acc.detach().cpu().numpy(), a0, atol=2e-5)
        g_rgb = rng.normal(size=(B, 3)).astype(np.float32)
        g_d = rng.normal(size=(B,)).astype(np.float32)
        (rgb * torch.from_numpy(g_rgb).cuda()).sum().add((depth * torch.from_numpy(g_d).cuda()).sum()).backward()
        ds0, dc0 = orc.volume_rendering_backward(sig, col, z, g_rgb, g_d, white)
        np.testing.assert_allclose(c_t.grad.cpu().numpy(), dc0, atol=2e-5)
        np.testing.assert_allclose(s_t.grad.cpu().numpy().reshape(B, N), ds0, atol=2e-5 * max(1.0, np.abs(ds0).max()))


@pytest.mark.gpu
def test_empty_inputs():
    import codenerf_b200 as cn
    model, _ = U.make_model("fp32")
    sc = torch.zeros(1, 256, device="cuda")
    sig, col = model(torch.zeros(0, 64, 3, device="cuda"), torch.zeros(0, 64, 3, device="cuda"), sc, sc)
    assert sig.shape == (0, 64, 1) and col.shape == (0, 64, 3)
    rgb, depth = cn.volume_rendering(sig, col, torch.linspace(0.8, 1.8, 64).cuda())
    assert rgb.shape == (0, 3) and depth.shape == (0,)


def _fused_vs_oracle(precision, atol, gtol, N, H, W, n_seg, ray_count, cat):
    """Camera-mode fused render of `n_seg` views' pixel windows vs the oracle."""
    import codenerf_b200 as cn
    model, flat = U.make_model(precision)
    focal = 131.25 * W / 128.0
    c2ws = np.stack([syn.look_at_pose(700 + g, cat["radius"]) for g in range(n_seg)])
    zs = np.stack([orc.z_vals(cat["near"], cat["far"], N, orc.torch_rand(4000 + g, N)) for g in range(n_seg)])
    pix = np.array([(37 * g) % (H * W - ray_count + 1) for g in range(n_seg)], np.int32)
    scodes, tcodes = syn.make_codes(11, n_seg), syn.make_codes(12, n_seg)
    tgt = syn.make_targets(13, n_seg * ray_count)
    ref = [orc.render(flat, H, W, focal, c2ws[g], zs[g], scodes[g:g + 1], tcodes[g:g + 1], True,
                      ray_begin=int(pix[g]), ray_count=ray_count) for g in range(n_seg)]
    bundle = cn.RayBundle(z_vals=torch.from_numpy(zs).cuda(), rays_per_segment=ray_count,
                          c2w=torch.from_numpy(c2ws).cuda(), pix_begin=torch.from_numpy(pix).cuda(),
                          focal=torch.tensor([focal], dtype=torch.float64), H=H, W=W)
    sc = torch.from_numpy(scodes).cuda().requires_grad_()
    tc = torch.from_numpy(tcodes).cuda().requires_grad_()
    rgb, depth, acc = cn.render(model, bundle, sc, tc)
    rgb_ref = np.concatenate([r["rgb"] for r in ref])
    np.testing.assert_allclose(rgb.detach().cpu().numpy(), rgb_ref, atol=atol, rtol=0)
    np.testing.assert_allclose(depth.detach().cpu().numpy(), np.concatenate([r["depth"] for r in ref]), atol=atol)
    np.testing.assert_allclose(acc.detach().cpu().numpy(), np.concatenate([r["acc"] for r in ref]), atol=atol)
    # loss: per-segment mean L2 (trainer.py:75) summed over segments
    t = torch.from_numpy(tgt).cuda()
    loss = ((rgb - t) ** 2).reshape(n_seg, -1).mean(1).sum()
    loss.backward()
    dP_ref = np.zeros(flat.size, np.float32)
    ds_ref, dt_ref = [], []
    for g in range(n_seg):
        d_rgb = (2.0 * (ref[g]["rgb"] - tgt[g * ray_count:(g + 1) * ray_count]) / (3.0 * ray_count)).astype(np.float32)
        dP, dsc, dtc = orc.render_backward(flat, ref[g], zs[g], scodes[g:g + 1], tcodes[g:g + 1], d_rgb, None, True)
        dP_ref += dP
        ds_ref.append(dsc)
        dt_ref.append(dtc)
    ds_ref, dt_ref = np.concatenate(ds_ref), np.concatenate(dt_ref)
    got = U.named_grads_flat(model)
    assert U.rel_err(sc.grad.cpu().numpy(), ds_ref) < gtol, U.rel_err(sc.grad.cpu().numpy(), ds_ref)
    assert U.rel_err(tc.grad.cpu().numpy(), dt_ref) < gtol
    o = 0
    for key, shp in orc.param_shapes():
        n = int(np.prod(shp))
        e = U.rel_err(got[o:o + n], dP_ref[o:o + n])
        assert e < gtol, (key, e)
        o += n
    return model, bundle, sc, tc, t, dP_ref, ds_ref, dt_ref, rgb_ref


@pytest.mark.gpu
@pytest.mark.parametrize("N,H,W,n_seg,ray_count,cat", [
    (64, 32, 32, 3, 128, syn.SRN_CARS),
    (96, 24, 40, 2, 100, syn.SRN_CHAIRS),
    (40, 16, 16, 1, 77, syn.SRN_CARS),
])
def test_fused_render_fp32_vs_oracle(N, H, W, n_seg, ray_count, cat):
    _fused_vs_oracle("fp32", ATOL, 2e-3, N, H, W, n_seg, ray_count, cat)


@pytest.mark.gpu
def test_train_step_fp32_matches_autograd_path():
    """cnb_render_train_step (fused loss seed) == render() + torch loss + backward."""
    from codenerf_b200 import ops, _lib
    N, H, W, n_seg, ray_count = 64, 32, 32, 3, 128
    model, bundle, sc, tc, t, dP_ref, ds_ref, dt_ref, rgb_ref = _fused_vs_oracle(
        "fp32", ATOL, 2e-3, N, H, W, n_seg, ray_count, syn.SRN_CARS)
    params = model.param_list()
    rb = bundle.args(sc.detach(), tc.detach())
    dP = torch.zeros(dP_ref.size, device="cuda")
    rgb, depth, acc, sq, dsc, dtc = ops.render_train_step(model._cfg, params, None, rb, _lib.PRECISION_FP32, t, 1.0, dP)
    np.testing.assert_allclose(rgb.cpu().numpy(), rgb_ref, atol=ATOL)
    sq_ref = ((rgb_ref - t.cpu().numpy()) ** 2).reshape(n_seg, -1).sum(1)
    np.testing.assert_allclose(sq.cpu().numpy(), sq_ref, rtol=1e-4)
    assert U.rel_err(dP.cpu().numpy(), dP_ref) < 2e-3
    assert U.rel_err(dsc.cpu().numpy(), ds_ref) < 2e-3
    assert U.rel_err(dtc.cpu().numpy(), dt_ref) < 2e-3
