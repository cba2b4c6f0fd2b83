"""Trainer / latent-fit loops on synthetic views: the bf16 tensor-core path against the
same loop in fp32 mode and against the reference loop restated on the CPU oracle.
North-star criterion: optimised-code PSNR within 0.05 dB."""
import math

import numpy as np
import pytest
import torch

from codenerf_b200 import synthetic as syn
from oracle import oracle as orc
from tests import gpu_util as U

HP = {"net_hyperparams": dict(syn.SRN_NET), "N_samples": 64, "near": 0.8, "far": 1.8, "loss_reg_coef": 1e-4,
      "lr_schedule": [{"type": "step", "lr": 1e-4, "interval": 250000}, {"type": "step", "lr": 1e-3, "interval": 250000}]}


def _targets_from_hidden_code(flat, H, W, focal, poses, seed):
    """Render target images with the oracle from a hidden code pair (so the fit has something to find)."""
    sc, tc = syn.make_codes(seed, 1) * 1.5, syn.make_codes(seed + 1, 1) * 1.5
    z = orc.z_vals(0.8, 1.8, 64, z_fixed=True)
    imgs = [orc.render(flat, H, W, focal, p, z, sc, tc, True)["rgb"] for p in poses]
    return np.stack(imgs).astype(np.float32)


def _fit(precision, flat_views, H, W, focal, poses, imgs, steps, seed):
    from codenerf_b200.optimizer import CodeFitter
    model, _ = U.make_model(precision)
    fitter = CodeFitter(model, HP, batch_size=H * W // 2, num_opts=steps)
    torch.manual_seed(seed)
    mean_s = torch.from_numpy(syn.make_codes(77, 1)[0])
    mean_t = torch.from_numpy(syn.make_codes(78, 1)[0])
    sc, tc, hist = fitter.fit(torch.tensor([focal], dtype=torch.float64), H, W, torch.from_numpy(imgs[:2]),
                              torch.from_numpy(poses[:2]), mean_s, mean_t, lr=1e-2, lr_half_interval=10)
    torch.manual_seed(seed + 1)
    ev = fitter.evaluate(torch.tensor([focal], dtype=torch.float64), H, W, torch.from_numpy(imgs[2:]),
                         torch.from_numpy(poses[2:]), sc, tc)
    return sc.cpu().numpy(), tc.cpu().numpy(), hist, ev


def _oracle_fit(flat, H, W, focal, poses, imgs, steps, seed):
    """The reference loop (src/optimizer.py:59-105) on the CPU oracle, AdamW from torch on CPU tensors."""
    B = H * W // 2
    torch.manual_seed(seed)
    sc = torch.from_numpy(syn.make_codes(77, 1)).clone().requires_grad_()
    tc = torch.from_numpy(syn.make_codes(78, 1)).clone().requires_grad_()
    nopts, hist = 0, []
    mk = lambda: torch.optim.AdamW([{"params": sc, "lr": 1e-2 * 2 ** (-(nopts // 10))}, {"params": tc, "lr": 1e-2 * 2 ** (-(nopts // 10))}])
    opt = mk()
    while nopts < steps:
        opt.zero_grad()
        gs, gt = np.zeros((1, 256), np.float32), np.zeros((1, 256), np.float32)
        for v in range(2):
            dist = (1.8 - 0.8) / (2 * 64)
            z = torch.linspace(0.8 + dist, 1.8 - dist, 64)
            z += torch.rand(64) * (1.8 - 0.8) / (2 * 64)
            z = z.numpy()
            fwd = orc.render(flat, H, W, focal, poses[v], z, sc.detach().numpy(), tc.detach().numpy(), True)
            d_rgb = (2.0 * (fwd["rgb"] - imgs[v]) / (3.0 * B)).astype(np.float32)
            _, dsc, dtc = orc.render_backward(flat, fwd, z, sc.detach().numpy(), tc.detach().numpy(), d_rgb, None, True,
                                              want_param_grads=False)
            s_n, t_n = sc.detach().numpy(), tc.detach().numpy()
            gs += dsc + 1e-4 * s_n / np.linalg.norm(s_n)
            gt += dtc + 1e-4 * t_n / np.linalg.norm(t_n)
            mses = ((fwd["rgb"] - imgs[v]) ** 2).reshape(-1, B * 3).mean(1)
        sc.grad, tc.grad = torch.from_numpy(gs), torch.from_numpy(gt)
        opt.step()
        hist.append(-10 * math.log(mses.mean()) / math.log(10))
        nopts += 1
        if nopts % 10 == 0:
            opt = mk()
    return sc.detach().numpy(), tc.detach().numpy(), hist


@pytest.mark.gpu
def test_latent_fit_psnr_parity():
    H = W = 16
    focal = 131.25 * W / 128
    flat, _ = syn.make_params(0)
    poses = np.stack([syn.look_at_pose(300 + i, 1.3) for i in range(4)])
    imgs = _targets_from_hidden_code(flat, H, W, focal, poses, 901)
    steps = 30
    s_o, t_o, h_o = _oracle_fit(flat, H, W, focal, poses, imgs, steps, 5)
    s_b, t_b, h_b, ev_b = _fit("bf16", flat, H, W, focal, poses, imgs, steps, 5)
    s_f, t_f, h_f, ev_f = _fit("fp32", flat, H, W, focal, poses, imgs, steps, 5)
    print("fit PSNR  oracle %.3f  fp32 %.3f  bf16 %.3f   (start %.3f)" % (h_o[-1], h_f[-1], h_b[-1], h_o[0]))
    print("eval PSNR fp32", ["%.3f" % v for v in ev_f], " bf16", ["%.3f" % v for v in ev_b])
    assert h_o[-1] > h_o[0] + 0.5                      # the fit actually improves
    assert abs(h_f[-1] - h_o[-1]) < 0.05               # fp32 path == reference loop
    assert abs(h_b[-1] - h_o[-1]) < 0.05               # bf16 path within 0.05 dB (north star)
    for a, b in zip(ev_b, ev_f):
        assert abs(a - b) < 0.05


@pytest.mark.gpu
def test_trainer_step_matches_fp32_loop():
    """A few AdamW iterations of the trainer idiom: bf16 and fp32 modes track each other."""
    from codenerf_b200.trainer import Trainer
    H = W = 32
    focal = torch.tensor([131.25 * W / 128], dtype=torch.float64)
    poses = torch.from_numpy(np.stack([syn.look_at_pose(400 + i, 1.3) for i in range(2)]))
    imgs = torch.from_numpy(np.stack([syn.make_targets(500 + i, H * W) for i in range(2)]))
    losses = {}
    for prec in ("fp32", "bf16"):
        torch.manual_seed(3)
        tr = Trainer(HP, n_objects=3, batch_size=512, precision=prec)
        flat, views = syn.make_params(0)
        tr.model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in views.items()})
        torch.manual_seed(4)
        ls = []
        for it in range(6):
            ls.append(float(tr.train_view(focal, H, W, imgs, poses, obj_idx=it % 3)))
        losses[prec] = ls
        st = tr.state()
        assert set(st) == {"model_params", "shape_code_params", "texture_code_params", "niter", "nepoch"}
        assert st["niter"] == 6
    print(losses)
    for a, b in zip(losses["fp32"], losses["bf16"]):
        assert abs(a - b) < 2e-3 * max(1.0, abs(a))
    assert losses["fp32"][-1] < losses["fp32"][0]
