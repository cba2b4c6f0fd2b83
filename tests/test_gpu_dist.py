"""The sharded drivers on two GPUs (NCCL, one process per GPU): SURVEY.md 8(e).  Skipped on a single-GPU box; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`.

* rendering shards (object, view) pairs, no collective: the gathered images equal the single-GPU images bit for bit;
* latent fitting shards objects, no collective: the gathered codes equal the single-GPU fit;
* training shards the objects of a step and all-reduces the flat MLP gradient: gradient and updated weights equal the
  single-GPU step on the concatenated batch.
"""
import os
import socket

import numpy as np
import pytest
import torch

from codenerf_b200 import synthetic as syn

HP = {"net_hyperparams": dict(syn.SRN_NET), "N_samples": 64, "near": 0.8, "far": 1.8, "loss_reg_coef": 1e-4,
      "lr_schedule": [{"type": "step", "lr": 1e-4, "interval": 250000}, {"type": "step", "lr": 1e-3, "interval": 250000}]}
H = W = 32
B = 512
FOCAL = 131.25 * W / 128
N_OBJ, N_VIEWS = 4, 3


def _inputs():
    poses = torch.from_numpy(np.stack([[syn.look_at_pose(40 + 9 * o + v, 1.3) for v in range(N_VIEWS)] for o in range(N_OBJ)]))
    imgs = torch.from_numpy(syn.make_targets(8, N_OBJ * N_VIEWS * H * W).reshape(N_OBJ, N_VIEWS, H * W, 3))
    sc, tc = torch.from_numpy(syn.make_codes(5, N_OBJ)), torch.from_numpy(syn.make_codes(6, N_OBJ))
    g = torch.Generator().manual_seed(9)
    zs = (torch.linspace(0.8, 1.8, 64)[None, None] + 0.004 * torch.rand(N_OBJ, N_VIEWS, 64, generator=g)).contiguous()
    return poses, imgs, sc, tc, zs


def _model():
    import codenerf_b200 as cn
    flat, views = syn.make_params(0)
    m = cn.CodeNeRF(**syn.SRN_NET, precision="bf16")
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in views.items()})
    return m.cuda()


def _run_all(rank_objs=None):
    """The three drivers in the current process (single GPU, or one rank of a group)."""
    from codenerf_b200.optimizer import CodeFitter, render_dataset
    from codenerf_b200.trainer import Trainer
    from codenerf_b200 import parallel
    poses, imgs, sc, tc, zs = _inputs()
    out = {}
    model = _model()
    r = render_dataset(model, HP, FOCAL, H, W, poses, sc, tc, targets=imgs, batch_size=B, views_per_launch=2, z_vals=zs,
                       keep_images=True)
    out["rgb"], out["psnr"] = r["rgb"].cpu(), r["psnr"].cpu()
    steps = 4
    fitter = CodeFitter(model, HP, batch_size=B, num_opts=steps)
    zf = zs[:, :2].permute(1, 0, 2)[None].expand(steps, 2, N_OBJ, 64).contiguous()
    s, t, h = fitter.fit_batch(FOCAL, H, W, imgs[:, :2], poses[:, :2], sc.mean(0), tc.mean(0), lr=1e-2, lr_half_interval=2, z_vals=zf)
    out["fit_s"], out["fit_t"], out["fit_h"] = s.cpu(), t.cpu(), h.cpu()
    # one training step: this rank's share of the 4 objects
    torch.manual_seed(3)
    tr = Trainer(HP, n_objects=N_OBJ, batch_size=B, precision="bf16")
    flat, views = syn.make_params(0)
    tr.model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in views.items()})
    rank, ws = parallel.world()
    b, e = parallel.shard_range(N_OBJ, ws, rank)
    objs = list(range(b, e))
    loss = tr.train_batch(FOCAL, H, W, imgs[objs, 0], poses[objs, 0], objs, z_vals=zs[objs, 0])
    out["grad"] = tr._dP.cpu()
    out["params"] = torch.cat([p.detach().reshape(-1) for p in tr.model.parameters()]).cpu()
    tr._owner = torch.tensor([0 if o < (N_OBJ + ws - 1) // ws or ws == 1 else 1 for o in range(N_OBJ)], dtype=torch.int32)
    st = tr.state()
    out["codes"] = st["shape_code_params"]["weight"].cpu()
    return out


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    import torch.distributed as dist
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    out = _run_all()
    if rank == 0:
        ret.update({k: v.numpy() for k, v in out.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_drivers_equal_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    single = {k: v.numpy() for k, v in _run_all().items()}
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    two = dict(ret)
    assert np.array_equal(two["rgb"], single["rgb"])                    # sharded render: bit for bit
    np.testing.assert_allclose(two["psnr"], single["psnr"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(two["fit_s"], single["fit_s"], rtol=0, atol=2e-4)     # atomics order inside a code's column sums
    np.testing.assert_allclose(two["fit_t"], single["fit_t"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(two["fit_h"], single["fit_h"], rtol=0, atol=1e-3)
    g1, g2 = single["grad"], two["grad"]
    assert np.abs(g2 - g1).max() < 1e-4 * np.abs(g1).max()              # all-reduced gradient == gradient of the whole batch
    assert np.abs(two["params"] - single["params"]).max() < 2.1e-4      # one AdamW step of lr 1e-4 (sign flips of ~0 gradients)
    assert np.abs(two["params"] - single["params"]).mean() < 1e-6
    np.testing.assert_allclose(two["codes"], single["codes"], rtol=0, atol=2.1e-3)
    assert np.abs(two["codes"] - single["codes"]).mean() < 1e-5
