"""SRN reader on a synthetic SRN-format tree; cross-checked against the reference's reader when the
reference checkout is present (authoring container only; imageio is stubbed with Pillow, which is what it wraps)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from codenerf_b200 import data as cd  # noqa: E402
from codenerf_b200 import synthetic as syn  # noqa: E402


def _make_tree(root, split, n_obj, n_views, H=128, W=128, focal=131.25):
    from PIL import Image
    base = os.path.join(root, "srn_cars", split)
    rng = np.random.RandomState(7)
    for o in range(n_obj):
        d = os.path.join(base, f"obj{o:03d}")
        os.makedirs(os.path.join(d, "rgb")); os.makedirs(os.path.join(d, "pose"))
        with open(os.path.join(d, "intrinsics.txt"), "w") as f:
            f.write(f"{focal} {W / 2} {H / 2} 0.\n0. 0. 0.\n1.\n{H} {W}\n")
        for v in range(n_views):
            img = rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
            Image.fromarray(img).save(os.path.join(d, "rgb", f"{v:06d}.png"))
            pose = syn.look_at_pose(100 * o + v, 1.3).astype(np.float64) @ np.diag([1.0, -1.0, -1.0, 1.0])   # stored in SRN's frame
            np.savetxt(os.path.join(d, "pose", f"{v:06d}.txt"), pose.reshape(1, 16))
    return base


def test_train_item_matches_the_file_contents(tmp_path):
    from PIL import Image
    _make_tree(str(tmp_path), "cars_train", 2, 50)
    ds = cd.SRN(data_dir=str(tmp_path), num_instances_per_obj=2, crop_img=True)
    assert len(ds) == 2 and ds.train
    np.random.seed(3)
    focal, H, W, imgs, poses, instances, idx = ds[1]
    assert (focal, H, W) == (131.25, 64, 64) and imgs.shape == (2, 64 * 64, 3) and poses.shape == (2, 4, 4)
    for k, v in enumerate(instances):
        raw = np.asarray(Image.open(tmp_path / "srn_cars" / "cars_train" / "obj001" / "rgb" / f"{v:06d}.png"), np.float32) / 255.0
        np.testing.assert_array_equal(imgs[k].numpy().reshape(64, 64, 3), raw[32:-32, 32:-32])
        np.testing.assert_allclose(poses[k].numpy(), syn.look_at_pose(100 + int(v), 1.3), atol=1e-6)   # the flip is undone
    # the decoded bytes are cached: a second read does not touch the files
    os.rename(tmp_path / "srn_cars" / "cars_train" / "obj001" / "rgb", tmp_path / "moved")
    again = ds.object(1).images(instances, crop=True)
    assert torch.equal(again.reshape(2, -1, 3), imgs)


def test_test_split_returns_all_views_uncropped(tmp_path):
    _make_tree(str(tmp_path), "cars_test", 1, 250, H=16, W=16)
    ds = cd.SRN(splits="cars_test", data_dir=str(tmp_path))
    focal, H, W, imgs, poses, idx = ds[0]
    assert not ds.train and (H, W) == (16, 16) and imgs.shape == (250, 16, 16, 3) and poses.shape == (250, 4, 4)


@pytest.mark.skipif(not os.path.exists("/root/reference/src/data.py"), reason="reference checkout not present")
def test_same_tensors_as_the_reference_reader(tmp_path):
    from PIL import Image
    _make_tree(str(tmp_path), "cars_train", 2, 50)
    stub = types.ModuleType("imageio")
    stub.imread = lambda path, pilmode="RGB": np.asarray(Image.open(path).convert(pilmode))
    saved = sys.modules.get("imageio")
    sys.modules["imageio"] = stub
    sys.path.insert(0, "/root/reference/src")
    try:
        import importlib
        ref_data = importlib.import_module("data")
    finally:
        sys.path.remove("/root/reference/src")
        if saved is None: sys.modules.pop("imageio", None)
        else: sys.modules["imageio"] = saved
    ref = ref_data.SRN(data_dir=str(tmp_path) + "/", num_instances_per_obj=3, crop_img=True)
    mine = cd.SRN(data_dir=str(tmp_path), num_instances_per_obj=3, crop_img=True)
    for idx in (0, 1):
        np.random.seed(11 + idx); a = ref[idx]
        np.random.seed(11 + idx); b = mine[idx]
        assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2] and a[6] == b[6]
        assert torch.equal(a[3], b[3]) and torch.equal(a[4], b[4]) and np.array_equal(a[5], b[5])


@pytest.mark.gpu
def test_train_batch_on_device_feeds_the_fused_step(tmp_path):
    """Objects decoded once into a device-resident uint8 cache; a batch of views goes straight into one fused launch."""
    import codenerf_b200 as cn
    from tests import gpu_util as U
    _make_tree(str(tmp_path), "cars_train", 3, 50)
    ds = cd.SRN(data_dir=str(tmp_path), crop_img=True, cache_device="cuda")
    rng = np.random.RandomState(5)
    focal, H, W, imgs, poses, views = ds.train_batch([0, 2, 1], device="cuda", rng=rng)
    assert imgs.is_cuda and imgs.shape == (3, 64 * 64, 3) and poses.shape == (3, 4, 4) and (H, W) == (64, 64)
    cpu = cd.SRN(data_dir=str(tmp_path), crop_img=True)
    for k, (o, v) in enumerate(zip([0, 2, 1], views)):
        assert torch.equal(imgs[k].cpu(), cpu.object(o).images([v], crop=True).reshape(-1, 3))
    model, flat = U.make_model("bf16")
    z = cn.make_z_vals(syn.SRN_CARS["near"], syn.SRN_CARS["far"], 64).cuda().reshape(1, -1).expand(3, -1).contiguous()
    bundle = cn.RayBundle(z_vals=z, rays_per_segment=H * W, c2w=poses, pix_begin=torch.zeros(3, dtype=torch.int32, device="cuda"),
                          focal=torch.tensor([focal], dtype=torch.float64), H=H, W=W)
    sc = torch.from_numpy(syn.make_codes(1, 3)).cuda().requires_grad_(); tc = torch.from_numpy(syn.make_codes(2, 3)).cuda().requires_grad_()
    rgb, depth, acc = cn.render(model, bundle, sc, tc)
    loss = torch.mean((rgb - imgs.reshape(-1, 3)) ** 2)
    loss.backward()
    assert torch.isfinite(loss) and torch.isfinite(sc.grad).all() and float(sc.grad.abs().max()) > 0
