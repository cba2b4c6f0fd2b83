"""SRN (ShapeNet cars / chairs) object reader feeding the fused render path.

Same on-disk format and the same view / crop conventions as reference src/data.py:10-88 --

    <data_dir>/<cat>/<split>/<object id>/{rgb/*.png, pose/*.txt, intrinsics.txt}

* pose files: 16 numbers, a row-major 4x4 camera-to-world matrix, right-multiplied by diag(1,-1,-1,1)
  (data.py:13-18: SRN's camera looks down +z with y down);
* images: 8-bit RGB, scaled by 1/255 (data.py:20-29);
* intrinsics.txt: focal = first token of the first line, `H W` = the last line (data.py:31-37);
* training objects: `num_instances_per_obj` views drawn with `np.random.choice(50, n)` from the first 50,
  centre crop `[32:-32, 32:-32]` with H, W halved when `crop_img` (data.py:62-74) -- the focal is NOT changed,
  exactly as in the reference; test / val objects: views 0..249 uncropped (data.py:76-88);

-- but organised for a GPU that renders millions of rays per second: every object is decoded ONCE into a uint8
cache (pinned host memory, or device memory with `cache_device="cuda"`), so a training step copies or indexes
bytes instead of decoding PNGs, and a whole batch of objects comes back as the tensors the fused op takes
(`poses [n, 4, 4]`, `imgs [n, H*W, 3]` fp32).  Files are listed in sorted order like the reference's `np.sort`.
"""
import os

import numpy as np
import torch

_SRN_FLIP = np.diag(np.array([1.0, -1.0, -1.0, 1.0]))
# float32(i) / float32(255) exactly as numpy computes it on the host (data.py:26-27); a table look-up gives the same
# bits on any device (a GPU division kernel is not guaranteed to round like the CPU's)
_U8_TO_UNIT = np.arange(256, dtype=np.float32) / np.float32(255.0)
_LUT = {}


def _unit_lut(device):
    key = str(device)
    if key not in _LUT:
        _LUT[key] = torch.from_numpy(_U8_TO_UNIT.copy()).to(device)
    return _LUT[key]


def read_intrinsics(path):
    """(focal, H, W) -- data.py:31-37."""
    with open(path, "r") as f:
        lines = f.readlines()
    focal = float(lines[0].split()[0])
    H, W = (int(t) for t in lines[-1].split())
    return focal, H, W


def read_pose(path):
    """camera-to-world [4, 4] float64 in the reference's convention (data.py:13-18)."""
    return np.loadtxt(path).reshape(4, 4) @ _SRN_FLIP


def _read_rgb8(path):
    from PIL import Image          # Pillow is what imageio's `pilmode='RGB'` uses underneath
    with Image.open(path) as im:
        return np.asarray(im.convert("RGB"), dtype=np.uint8)


class SRNObject:
    """One object directory, decoded lazily and cached as uint8."""

    def __init__(self, obj_dir, cache_device="cpu"):
        self.dir = obj_dir
        self.cache_device = torch.device(cache_device)
        self.focal, self.H, self.W = read_intrinsics(os.path.join(obj_dir, "intrinsics.txt"))
        self.pose_files = sorted(os.path.join(obj_dir, "pose", f.name) for f in os.scandir(os.path.join(obj_dir, "pose")))
        self.img_files = sorted(os.path.join(obj_dir, "rgb", f.name) for f in os.scandir(os.path.join(obj_dir, "rgb")))
        self._poses = None
        self._imgs = {}              # view index -> uint8 [H, W, 3] tensor on cache_device (pinned when on the host)

    @property
    def n_views(self):
        return len(self.img_files)

    def poses(self, idxs):
        if self._poses is None:
            self._poses = torch.from_numpy(np.stack([read_pose(p) for p in self.pose_files])).float()
        return self._poses[torch.as_tensor(np.asarray(idxs), dtype=torch.long)]

    def image_u8(self, i):
        i = int(i)
        t = self._imgs.get(i)
        if t is None:
            t = torch.from_numpy(_read_rgb8(self.img_files[i]).copy())
            if self.cache_device.type == "cuda":
                t = t.to(self.cache_device)
            elif torch.cuda.is_available():
                t = t.pin_memory()
            self._imgs[i] = t
        return t

    def images(self, idxs, crop=False, device=None):
        """[n, H', W', 3] fp32 in [0, 1] (uint8 / 255, like data.py:26-27), optionally centre-cropped."""
        dev = torch.device(device) if device is not None else self.cache_device
        out = []
        for i in idxs:
            u8 = self.image_u8(i)
            if crop:
                u8 = u8[32:-32, 32:-32, :]
            out.append(_unit_lut(dev)[u8.to(dev, non_blocking=True).long()])
        return torch.stack(out)


class SRN:
    """Dataset view with the reference's `__getitem__` contract (data.py:39-88)."""

    def __init__(self, cat="srn_cars", splits="cars_train", data_dir="../data/ShapeNet_SRN/", num_instances_per_obj=1,
                 crop_img=True, cache_device="cpu"):
        self.data_dir = os.path.join(data_dir, cat, splits)
        self.ids = sorted(f.name for f in os.scandir(self.data_dir))
        self.lenids = len(self.ids)
        self.num_instances_per_obj = num_instances_per_obj
        self.train = splits.split("_")[1] == "train"
        self.crop_img = crop_img
        self.cache_device = cache_device
        self._objects = {}

    def __len__(self):
        return self.lenids

    def object(self, idx):
        o = self._objects.get(idx)
        if o is None:
            o = self._objects[idx] = SRNObject(os.path.join(self.data_dir, self.ids[idx]), self.cache_device)
        return o

    def __getitem__(self, idx):
        o = self.object(idx)
        if self.train:
            instances = np.random.choice(50, self.num_instances_per_obj)          # data.py:66
            imgs = o.images(instances, crop=self.crop_img)
            H, W = (o.H // 2, o.W // 2) if self.crop_img else (o.H, o.W)
            return o.focal, H, W, imgs.reshape(self.num_instances_per_obj, -1, 3), o.poses(instances), instances, idx
        instances = np.arange(250)                                                # data.py:80
        return o.focal, o.H, o.W, o.images(instances), o.poses(instances), idx

    def train_batch(self, obj_indices, device="cuda", rng=None):
        """Several training objects at once, one view each, as device tensors for one fused launch:
        (focal, H, W, imgs [n, H*W, 3], poses [n, 4, 4], view index per object)."""
        rng = rng if rng is not None else np.random
        imgs, poses, views = [], [], []
        focal = H = W = None
        for idx in obj_indices:
            o = self.object(int(idx))
            v = int(rng.choice(50, 1)[0])
            im = o.images([v], crop=self.crop_img, device=device)
            h, w = im.shape[1], im.shape[2]
            if focal is None:
                focal, H, W = o.focal, h, w
            elif (o.focal, h, w) != (focal, H, W):
                raise ValueError("objects of one batch must share intrinsics")
            imgs.append(im.reshape(-1, 3)); poses.append(o.poses([v])[0]); views.append(v)
        return focal, H, W, torch.stack(imgs), torch.stack(poses).to(device), views
