"""The UNMODIFIED reference (yuliangguo/code-nerf src/model.py + src/utils.py) as a timed baseline.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/oracle.py): imported by bench.py's `cpu_baseline` and
`--impl reference` legs and by tests; the product package never imports it.

`/root/reference` exists only in the authoring container.  `stage()` -- called by
`__graft_entry__.build()` there -- copies the two files the render path consists of, byte for byte,
into `oracle/_ref/` (git-ignored build output, like a compiled `.so`: it travels to the GPU box with the
tree but never enters the history).  Nothing here reads `/root/reference` at run time.

`imageio` (imported by utils.py:2, unused by the path) is not installed anywhere: it is stubbed with
an empty module before the import, exactly as tests/golden/make_golden.py does.
"""
import importlib.util
import os
import shutil
import sys
import time
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
FILES = ("model.py", "utils.py")


def stage(src="/root/reference/src"):
    """Copy the reference's two hot-path files into oracle/_ref/ (no-op when the checkout is absent)."""
    if not all(os.path.exists(os.path.join(src, f)) for f in FILES):
        return False
    os.makedirs(REF_DIR, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(REF_DIR, f))
    return True


def available():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in FILES)


_mods = None


def load():
    """(utils module, model module) of the staged reference, or None when it is not staged."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        return None
    sys.modules.setdefault("imageio", types.ModuleType("imageio"))
    out = []
    for name in ("utils", "model"):
        spec = importlib.util.spec_from_file_location(f"codenerf_reference_{name}", os.path.join(REF_DIR, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        out.append(m)
    _mods = tuple(out)
    return _mods


class TrainChunk:
    """The body of the reference's training iteration (src/trainer.py:65-82) for one view and ONE 2048-ray chunk of it,
    written with the reference's own functions and call order: get_rays -> sample_from_rays (whole view, on the host,
    as the reference does) -> slice the chunk and move it to the device -> CodeNeRF.forward -> volume_rendering ->
    L2 loss + code regulariser -> backward."""

    def __init__(self, device, n_samples, state_dict, shape_code, texture_code, c2w, target, H=128, W=128, focal=131.25,
                 near=0.8, far=1.8, batch=2048, net=None):
        import torch
        self.torch = torch
        self.utils, model_mod = load()
        self.dev = torch.device(device)
        self.model = model_mod.CodeNeRF(**(net or {})).to(self.dev)
        self.model.load_state_dict(state_dict)
        self.shape = torch.as_tensor(shape_code).reshape(1, -1).clone().to(self.dev).requires_grad_()
        self.tex = torch.as_tensor(texture_code).reshape(1, -1).clone().to(self.dev).requires_grad_()
        self.c2w = torch.as_tensor(c2w).float()
        self.target = torch.as_tensor(target).float()[:batch]
        self.H, self.W, self.N, self.B = H, W, n_samples, batch
        self.focal = torch.tensor([focal], dtype=torch.float64)       # what DataLoader collation yields (src/data.py:34)
        self.near, self.far = near, far

    def raygen(self):
        ro, vd = self.utils.get_rays(self.H, self.W, self.focal, self.c2w)
        return self.utils.sample_from_rays(ro, vd, self.near, self.far, self.N)

    def chunk(self, xyz, viewdir, z_vals, backward=True):
        torch = self.torch
        B = self.B
        with torch.set_grad_enabled(backward):
            sigmas, rgbs = self.model(xyz[:B].to(self.dev), viewdir[:B].to(self.dev), self.shape, self.tex)
            rgb_rays, _ = self.utils.volume_rendering(sigmas, rgbs, z_vals.to(self.dev))
            loss = torch.mean((rgb_rays - self.target.type_as(rgb_rays).to(self.dev)) ** 2)
            if backward:
                reg = torch.norm(self.shape, dim=-1) + torch.norm(self.tex, dim=-1)
                (loss + 1e-4 * torch.mean(reg)).backward()
        return float(loss.item())          # the reference syncs on .item() every chunk (trainer.py:83)

    def timed(self, steps, warmup, backward=True):
        """Seconds per view-raygen and per chunk (mean over `steps`).  rays/s of a whole-view iteration with n chunks
        = n * B / (t_raygen + n * t_chunk)."""
        torch = self.torch
        sync = (lambda: torch.cuda.synchronize(self.dev)) if self.dev.type == "cuda" else (lambda: None)
        for _ in range(warmup):
            args = self.raygen()
            self.chunk(*args, backward=backward)
            self.model.zero_grad(set_to_none=True)
        sync()
        t_ray = t_chunk = 0.0
        for _ in range(steps):
            t0 = time.perf_counter()
            args = self.raygen()
            t1 = time.perf_counter()
            self.chunk(*args, backward=backward)
            sync()
            t2 = time.perf_counter()
            self.model.zero_grad(set_to_none=True)
            t_ray += t1 - t0
            t_chunk += t2 - t1
        return t_ray / steps, t_chunk / steps
