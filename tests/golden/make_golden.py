"""Generate the golden fixtures in tests/golden/ by executing the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference, which does not exist on the
GPU box):   python tests/golden/make_golden.py

The reference (yuliangguo/code-nerf) has no tests or golden vectors of its own, so these
fixtures -- outputs of its own src/utils.py and src/model.py on seeded synthetic inputs --
are what pins the oracle (oracle/codenerf_oracle.c) and, through it, the CUDA path.
Inputs come from codenerf_b200/synthetic.py (hash-based, machine independent), so a
fixture stores only outputs plus the few scalars that name the case.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("CODENERF_REFERENCE", "/root/reference")
sys.modules.setdefault("imageio", types.ModuleType("imageio"))   # utils.py:2 imports it; hot path never uses it
sys.path.insert(0, os.path.join(REF, "src"))
import utils as ref_utils      # noqa: E402  (reference src/utils.py)
import model as ref_model      # noqa: E402  (reference src/model.py)

from codenerf_b200 import synthetic as syn   # noqa: E402

torch.set_num_threads(1)       # fixed reduction order for the float fixtures


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def checksum(a):
    """Order-dependent 64-bit checksum of fp32 bit patterns."""
    b = bits(a).ravel().astype(np.uint64)
    idx = np.arange(b.size, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return np.uint64(np.bitwise_xor.reduce((b + np.uint64(1)) * (idx * np.uint64(2654435761) + np.uint64(97))))


def rays_cases():
    """get_rays + sample_from_rays, bit-exact fixtures."""
    out = {}
    cases = []
    k = 0
    for cat in (syn.SRN_CARS, syn.SRN_CHAIRS):
        for N in (64, 96):
            for (H, W) in ((128, 128), (64, 64), (16, 24)):
                for focal, f64 in ((131.25, True), (119.4716, True), (131.25, False), (98.76543, False)):
                    if (H, W) != (16, 24) and focal != 131.25:
                        continue
                    cases.append((k, cat, N, H, W, focal, f64))
                    k += 1
    meta = []
    for (k, cat, N, H, W, focal, f64) in cases:
        c2w = torch.from_numpy(syn.look_at_pose(100 + k, cat["radius"]))
        focal_arg = torch.tensor([focal], dtype=torch.float64) if f64 else focal   # data.py:34 collated -> fp64 [1]
        ro, vd = ref_utils.get_rays(H, W, focal_arg, c2w)
        torch.manual_seed(1000 + k)
        xyz, vdr, z = ref_utils.sample_from_rays(ro, vd, cat["near"], cat["far"], N)
        zf = ref_utils.sample_from_rays(ro[:1], vd[:1], cat["near"], cat["far"], N, z_fixed=True)[2]
        assert torch.equal(vdr[:, 0], vd)
        sub = np.arange(0, H * W, 389)
        out[f"c{k}_rays_o"] = bits(ro.numpy()[sub])
        out[f"c{k}_viewdirs_sub"] = bits(vd.numpy()[::13])
        out[f"c{k}_viewdirs_sum"] = np.array([checksum(vd.numpy())], dtype=np.uint64)
        if H * W <= 4096:
            out[f"c{k}_viewdirs"] = bits(vd.numpy())
        out[f"c{k}_z"] = bits(z.numpy())
        out[f"c{k}_zfixed"] = bits(zf.numpy())
        out[f"c{k}_xyz_sub"] = bits(xyz.numpy()[sub])
        out[f"c{k}_xyz_sum"] = np.array([checksum(xyz.numpy())], dtype=np.uint64)
        meta.append([k, cat["near"], cat["far"], cat["radius"], N, H, W, focal, 1.0 if f64 else 0.0])
    out["meta"] = np.array(meta, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "rays_samples.npz"), **out)
    print("rays_samples.npz:", len(cases), "cases")


def load_ref_model(cfg, params):
    m = ref_model.CodeNeRF(**cfg)
    sd = {k: torch.from_numpy(v.copy()) for k, v in params.items()}
    m.load_state_dict(sd)
    return m


def weight_probe(seed, n):
    return syn.uniform(seed + 31337, n, -1.0, 1.0)


def grad_summary(named_grads):
    """Compact fingerprint of a full gradient set: per tensor [sum, abs-sum, dot(probe)] in fp64,
    plus biases in full and the first 4 rows of each weight."""
    out = {}
    for t, (k, g) in enumerate(named_grads):
        g64 = g.astype(np.float64).ravel()
        out[f"stat/{k}"] = np.array([g64.sum(), np.abs(g64).sum(), float(g64 @ weight_probe(t, g64.size))])
        if g.ndim == 1:
            out[f"full/{k}"] = g.astype(np.float32)
        else:
            out[f"head/{k}"] = g[:4].astype(np.float32)
    return out


def render_cases():
    """PE, CodeNeRF.forward, volume_rendering and autograd fixtures on a small view."""
    out = {}
    meta = []
    specs = [
        # id, category, N, H, W, n_codes (1 = broadcast [1,256] as trainer.py:70), white_bg
        (0, syn.SRN_CARS, 64, 16, 16, 1, True),
        (1, syn.SRN_CHAIRS, 96, 12, 20, 1, True),
        (2, syn.SRN_CARS, 64, 16, 16, 4, True),      # per-ray codes [B,1,256]
        (3, syn.SRN_CARS, 40, 8, 8, 1, False),       # ragged N, black background
    ]
    cfg = dict(syn.SRN_NET)
    flat, params = syn.make_params(0, cfg)
    model = load_ref_model(cfg, params)
    for (k, cat, N, H, W, n_codes, white) in specs:
        R = H * W
        c2w = torch.from_numpy(syn.look_at_pose(200 + k, cat["radius"]))
        focal = torch.tensor([syn.SRN_FOCAL * W / 128.0], dtype=torch.float64)
        ro, vd = ref_utils.get_rays(H, W, focal, c2w)
        torch.manual_seed(2000 + k)
        xyz, vdr, z = ref_utils.sample_from_rays(ro, vd, cat["near"], cat["far"], N)
        sc = torch.from_numpy(syn.make_codes(300 + k, n_codes)).requires_grad_()
        tc = torch.from_numpy(syn.make_codes(400 + k, n_codes)).requires_grad_()
        if n_codes == 1:
            sc_in, tc_in = sc, tc                                       # [1,256]
        else:
            per = R // n_codes
            sc_in = sc.repeat_interleave(per, 0).unsqueeze(1)           # [R,1,256]
            tc_in = tc.repeat_interleave(per, 0).unsqueeze(1)
        model.zero_grad()
        sig, col = model(xyz, vdr, sc_in, tc_in)                        # model.py:36
        rgb, depth = ref_utils.volume_rendering(sig, col, z, white_bg=white)
        # accumulation = weights.sum(1) (utils.py:45), re-derived with the same torch ops
        with torch.no_grad():
            deltas = torch.cat([z[1:] - z[:-1], torch.ones(1) * 1e10])
            alphas = 1 - torch.exp(-sig.squeeze(-1) * deltas)
            trans = 1 - alphas + 1e-10
            T = torch.cumprod(torch.cat([torch.ones_like(trans[..., :1]), trans], -1), -1)[..., :-1]
            acc = (alphas * T).sum(1)
        tgt = torch.from_numpy(syn.make_targets(500 + k, R))
        loss_l2 = torch.mean((rgb - tgt) ** 2)                          # trainer.py:75
        reg = torch.norm(sc_in, dim=-1) + torch.norm(tc_in, dim=-1)     # trainer.py:77
        loss = loss_l2 + 1e-4 * torch.mean(reg) + 0.37 * depth.mean()   # depth term exercises d_depth
        loss.backward()
        pe_x = ref_model.PE(xyz[:3], 10)
        pe_d = ref_model.PE(vdr[:3], 4)
        out[f"r{k}_z"] = bits(z.numpy())
        out[f"r{k}_pe_xyz"] = pe_x.detach().numpy()
        out[f"r{k}_pe_dir"] = pe_d.detach().numpy()
        out[f"r{k}_sigmas"] = sig.detach().numpy().reshape(R, N)
        out[f"r{k}_rgbs"] = col.detach().numpy()
        out[f"r{k}_rgb"] = rgb.detach().numpy()
        out[f"r{k}_depth"] = depth.detach().numpy()
        out[f"r{k}_acc"] = acc.numpy()
        out[f"r{k}_loss"] = np.array([loss_l2.item(), loss.item()])
        out[f"r{k}_d_shape"] = sc.grad.numpy()
        out[f"r{k}_d_tex"] = tc.grad.numpy()
        named = [(key, p.grad.numpy()) for key, p in model.named_parameters()]
        for kk, v in grad_summary(named).items():
            out[f"r{k}_g/{kk}"] = v
        meta.append([k, cat["near"], cat["far"], cat["radius"], N, H, W, n_codes, 1.0 if white else 0.0])
    out["meta"] = np.array(meta, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "render_grads.npz"), **out)
    print("render_grads.npz:", len(specs), "cases")


if __name__ == "__main__":
    rays_cases()
    render_cases()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
