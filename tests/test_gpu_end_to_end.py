"""The rows of SURVEY.md 8(f) chained on the fused path, the way train.py -> optimize.py use them:
SRN tree on disk -> reader (f4) -> Trainer iterations (f1) -> models.pth (f3) -> CodeFitter from the checkpoint (f2).
The images are rendered by a 'teacher' network with hidden codes, so the losses have something to learn."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from codenerf_b200 import synthetic as syn  # noqa: E402
from tests import gpu_util as U  # noqa: E402

HPAMS = {"net_hyperparams": dict(syn.SRN_NET), "N_samples": 64, "near": syn.SRN_CARS["near"], "far": syn.SRN_CARS["far"],
         "loss_reg_coef": 1e-4,
         "lr_schedule": [{"type": "step", "lr": 1e-4, "interval": 250000}, {"type": "step", "lr": 1e-3, "interval": 250000}]}


def _teacher_tree(root, n_obj, n_views):
    """<root>/srn_cars/cars_train/<obj>/{rgb,pose,intrinsics.txt} with images rendered by a random-init teacher."""
    import codenerf_b200 as cn
    from PIL import Image
    teacher, _ = U.make_model("bf16")
    H = W = 128
    focal = syn.SRN_FOCAL
    base = os.path.join(root, "srn_cars", "cars_train")
    z = cn.make_z_vals(HPAMS["near"], HPAMS["far"], 64, z_fixed=True)
    for o in range(n_obj):
        d = os.path.join(base, f"{o:04d}")
        os.makedirs(os.path.join(d, "rgb")); os.makedirs(os.path.join(d, "pose"))
        with open(os.path.join(d, "intrinsics.txt"), "w") as f:
            f.write(f"{focal} {W / 2} {H / 2} 0.\n0. 0. 0.\n1.\n{H} {W}\n")
        sc = torch.from_numpy(syn.make_codes(500 + o, 1)).cuda() * 3.0
        tc = torch.from_numpy(syn.make_codes(600 + o, 1)).cuda() * 3.0
        for v in range(n_views):
            c2w = syn.look_at_pose(50 * o + v, syn.SRN_CARS["radius"])
            with torch.no_grad():
                rgb, _, _ = cn.render_view(teacher, H, W, torch.tensor([focal], dtype=torch.float64), torch.from_numpy(c2w), z, sc, tc)
            img = (rgb.clamp(0, 1) * 255.0 + 0.5).to(torch.uint8).reshape(H, W, 3).cpu().numpy()
            Image.fromarray(img).save(os.path.join(d, "rgb", f"{v:06d}.png"))
            np.savetxt(os.path.join(d, "pose", f"{v:06d}.txt"), (c2w.astype(np.float64) @ np.diag([1.0, -1.0, -1.0, 1.0])).reshape(1, 16))
    return base


@pytest.mark.gpu
def test_reader_trainer_checkpoint_fitter(tmp_path):
    from codenerf_b200 import checkpoint as ck
    from codenerf_b200 import data as cd
    from codenerf_b200.optimizer import CodeFitter
    from codenerf_b200.trainer import Trainer
    torch.manual_seed(0); np.random.seed(0)
    _teacher_tree(str(tmp_path), n_obj=2, n_views=50)
    ds = cd.SRN(data_dir=str(tmp_path), num_instances_per_obj=1, crop_img=True, cache_device="cuda")
    tr = Trainer(HPAMS, n_objects=len(ds), device="cuda", precision="bf16")
    losses = []
    for it in range(8):
        for obj in range(len(ds)):
            focal, H, W, imgs, poses, instances, idx = ds[obj]
            loss = tr.train_view(torch.tensor([focal], dtype=torch.float64), H, W, imgs, poses, idx)
            losses.append(float(loss))
    assert all(np.isfinite(losses)) and tr.niter == 16
    assert np.mean(losses[-4:]) < np.mean(losses[:4])                 # it learns something in 16 AdamW steps

    saved = tr.save_models(str(tmp_path / "exp"), iteration=16)
    assert tuple(saved.keys()) == ck.MODEL_KEYS and os.path.exists(tmp_path / "exp" / "16.pth")

    fitter, mean_s, mean_t = CodeFitter.from_checkpoint(str(tmp_path / "exp" / "models.pth"), HPAMS, device="cuda", num_opts=12)
    for (k1, v1), (k2, v2) in zip(tr.model.state_dict().items(), fitter.model.state_dict().items()):
        assert k1 == k2 and torch.equal(v1.cpu(), v2.cpu())
    assert torch.allclose(mean_s[0], tr.shape_codes.weight.detach().mean(0).cpu())
    o = ds.object(0)
    views = [3, 17]
    imgs = o.images(views, crop=True).reshape(2, -1, 3)
    s, t, hist = fitter.fit(torch.tensor([o.focal], dtype=torch.float64), 64, 64, imgs, o.poses(views), mean_s, mean_t, lr=1e-2, lr_half_interval=50)
    assert len(hist) == 12 and all(np.isfinite(hist)) and hist[-1] > hist[0]
    ev = fitter.evaluate(torch.tensor([o.focal], dtype=torch.float64), 64, 64, o.images([30], crop=True).reshape(1, -1, 3), o.poses([30]), s, t)
    assert np.isfinite(ev[0])
    ck.save_codes(str(tmp_path / "exp"), list(ds.ids), 0, s.cpu(), t.cpu(), {0: ev}, {})
    assert ck.load_codes(str(tmp_path / "exp" / "codes.pth"))["psnr_eval"][0] == ev
