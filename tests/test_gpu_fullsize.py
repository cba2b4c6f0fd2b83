"""BASELINE.json's full-size workload (32 objects x 2048 rays x 64 samples = 4 Mi samples per step) through
size-independent properties: the oracle cannot run this size in seconds, these checks can.

* a ray's colour does not depend on the batch it is rendered in (bit-exact: rows are independent in the MMAs);
* the fused training step reports the forward render's squared error;
* the backward pass is linear in its seed and additive over objects;
* the 1 Mi-row sub-batching of the training step does not change the result.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from codenerf_b200 import synthetic as syn  # noqa: E402
from tests import gpu_util as U  # noqa: E402

N_OBJ, RAYS, N = 32, 2048, 64


def _batch(n_obj=N_OBJ):
    import bench
    import codenerf_b200 as cn
    model, flat = U.make_model("bf16")
    c2w, pix, z, tgt, sc, tc = bench.synthetic_batch(n_obj, 0)
    dev = "cuda"
    focal = torch.tensor([syn.SRN_FOCAL], dtype=torch.float64)
    t = dict(c2w=torch.from_numpy(c2w).to(dev), pix=torch.from_numpy(pix).to(dev), z=torch.from_numpy(z).to(dev),
             tgt=torch.from_numpy(tgt).to(dev), sc=torch.from_numpy(sc).to(dev), tc=torch.from_numpy(tc).to(dev))

    def bundle(lo, hi):
        return cn.RayBundle(z_vals=t["z"][lo:hi], rays_per_segment=RAYS, c2w=t["c2w"][lo:hi], pix_begin=t["pix"][lo:hi],
                            focal=focal, H=syn.SRN_HW, W=syn.SRN_HW)
    return model, t, bundle


def _no_timeouts():
    from codenerf_b200 import _lib
    assert _lib.load().cnb_debug_pipeline_timeouts() == 0


@pytest.mark.gpu
def test_full_batch_render_equals_per_object_render():
    import codenerf_b200 as cn
    model, t, bundle = _batch()
    with torch.no_grad():
        rgb, depth, acc = cn.render(model, bundle(0, N_OBJ), t["sc"], t["tc"])
        for g in (0, 7, 19, 31):
            r1, d1, a1 = cn.render(model, bundle(g, g + 1), t["sc"][g:g + 1], t["tc"][g:g + 1])
            sl = slice(g * RAYS, (g + 1) * RAYS)
            assert torch.equal(rgb[sl], r1) and torch.equal(depth[sl], d1) and torch.equal(acc[sl], a1), g
    _no_timeouts()
    assert torch.isfinite(rgb).all() and float(acc.min()) >= 0.0 and float(acc.max()) <= 1.0 + 1e-5


@pytest.mark.gpu
def test_train_step_reports_the_forward_error_and_is_additive_over_objects():
    import codenerf_b200 as cn
    from codenerf_b200 import _lib, ops
    model, t, bundle = _batch()
    params = model.param_list(); packed = model._packed.get(model._cfg, params)
    prec = _lib.PRECISION_BF16
    n_par = sum(p.numel() for p in params)

    def step(lo, hi):
        dP = torch.zeros(n_par, device="cuda")
        rb = bundle(lo, hi).args(t["sc"][lo:hi], t["tc"][lo:hi])
        out = ops.render_train_step(model._cfg, params, packed, rb, prec, t["tgt"][lo * RAYS:hi * RAYS], 1.0, dP, want_outputs=True)
        return dP, out

    dP_all, out_all = step(0, N_OBJ)                    # 4 Mi rows: four 1 Mi-row sub-batches
    rgb, sq, dsc = out_all[0], out_all[3], out_all[4]
    _no_timeouts()
    # (1) the per-object squared error is the forward render's
    with torch.no_grad():
        rgb_f, _, _ = cn.render(model, bundle(0, N_OBJ), t["sc"], t["tc"])
    assert torch.equal(rgb, rgb_f)
    ref_sq = ((rgb_f - t["tgt"]) ** 2).reshape(N_OBJ, -1).double().sum(1)
    np.testing.assert_allclose(sq.double().cpu().numpy(), ref_sq.cpu().numpy(), rtol=1e-4)
    # (2) additivity: the gradient of the batch is the sum of the gradients of its halves (one sub-batch each would
    #     hide sub-batching bugs, so use unequal parts: 5 + 27 objects)
    dP_a, out_a = step(0, 5)
    dP_b, out_b = step(5, N_OBJ)
    tot = dP_a + dP_b
    scale = float(dP_all.abs().max())
    assert scale > 0
    assert float((dP_all - tot).abs().max()) <= 2e-3 * scale
    # code gradients are per object: identical rows whatever the batch (same tiles, same arithmetic, fp32 atomics aside)
    dsc_parts = torch.cat([out_a[4], out_b[4]])
    assert float((dsc - dsc_parts).abs().max()) <= 2e-3 * float(dsc.abs().max())


@pytest.mark.gpu
def test_backward_is_linear_in_the_seed():
    from codenerf_b200 import _lib, ops
    model, t, bundle = _batch(8)
    params = model.param_list(); packed = model._packed.get(model._cfg, params)
    rb = bundle(0, 8).args(t["sc"][:8], t["tc"][:8])
    seed = (t["tgt"][:8 * RAYS] - 0.5) * 1e-3
    g1 = ops.render_backward(model._cfg, params, packed, rb, _lib.PRECISION_BF16, seed, None, True)
    g2 = ops.render_backward(model._cfg, params, packed, rb, _lib.PRECISION_BF16, seed * 4.0, None, True)   # x4: exact in bf16 / fp32
    _no_timeouts()
    for a, b in zip(g1, g2):
        s = float(b.abs().max())
        assert s > 0 and float((a * 4.0 - b).abs().max()) <= 1e-4 * s


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{"CNB_BWD_PAIRS": "0"}, {"CNB_WEIGHT_MCAST": "2", "CNB_BWD_PAIRS": "0"}, {"CNB_K3_OVERLAP": "1"},
                                 {"CNB_STASH_LANES": "32"}],
                         ids=["single-cta", "multicast2", "k3-overlap", "stash-2KB-pieces"])
def test_backward_kernel_variants_match_default(env):
    """The other forms of the training step (measured, not faster than the default CTA-pair kernel) give
    the default's gradients."""
    from codenerf_b200 import _lib, ops
    model, t, bundle = _batch()
    params = model.param_list(); packed = model._packed.get(model._cfg, params)
    n_par = sum(p.numel() for p in params)

    def step():
        dP = torch.zeros(n_par, device="cuda")
        rb = bundle(0, N_OBJ).args(t["sc"], t["tc"])
        out = ops.render_train_step(model._cfg, params, packed, rb, _lib.PRECISION_BF16, t["tgt"], 1.0, dP, want_outputs=True)
        torch.cuda.synchronize()
        return [dP] + [o for o in out if o is not None]

    base = step()
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        got = step()
    finally:
        for k, v in old.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v
    _no_timeouts()
    for a, b in zip(got, base):
        s = float(b.abs().max())
        assert float((a - b).abs().max()) <= 1e-4 * max(s, 1e-30)
