"""N > 1 host logic on CPU (gloo, world_size 2): object sharding + the one gradient all-reduce give the
same MLP gradient as a single process over the whole batch.  The per-rank arithmetic is the oracle (no GPU
in this container); what is under test is codenerf_b200/parallel.py."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from codenerf_b200 import parallel
from codenerf_b200 import synthetic as syn
from oracle import oracle as orc

N, H, W, RAYS, N_OBJ = 16, 8, 8, 8, 3


def _object_grad(flat, g):
    cat = syn.SRN_CARS
    c2w = syn.look_at_pose(50 + g, cat["radius"])
    z = orc.z_vals(cat["near"], cat["far"], N, orc.torch_rand(60 + g, N))
    sc, tc = syn.make_codes(70 + g, 1), syn.make_codes(80 + g, 1)
    tgt = syn.make_targets(90 + g, RAYS)
    fwd = orc.render(flat, H, W, 131.25 * W / 128, c2w, z, sc, tc, True, ray_begin=0, ray_count=RAYS)
    d_rgb = (2.0 * (fwd["rgb"] - tgt) / (3.0 * RAYS)).astype(np.float32)
    dP, dsc, dtc = orc.render_backward(flat, fwd, z, sc, tc, d_rgb, None, True)
    return dP


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat, _ = syn.make_params(0)
    b, e = parallel.shard_range(N_OBJ, world, rank)
    g = np.zeros(flat.size, np.float32)
    for obj in range(b, e):
        g += _object_grad(flat, obj)
    t = torch.from_numpy(g)
    parallel.allreduce_mlp_grad(t)
    mx = parallel.max_over_ranks(float(rank + 1), torch.device("cpu"))
    if rank == 0:
        out["grad"] = t.numpy().copy()
        out["max"] = mx
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 64, 1024):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def test_two_rank_gradient_allreduce_equals_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.get_context("spawn").Manager()      # no fork() from a multi-threaded pytest process
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    flat, _ = syn.make_params(0)
    ref = np.zeros(flat.size, np.float32)
    for obj in range(N_OBJ):
        ref += _object_grad(flat, obj)
    np.testing.assert_allclose(out["grad"], ref, rtol=1e-5, atol=1e-9)
    assert out["max"] == 2.0


def _worker_tables(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank holds the full table; only the rows it owns are current (stale rows hold garbage)
    owner = torch.tensor([0, 1, 1, 0, 1])
    table = torch.full((5, 3), -99.0)
    for i in range(5):
        if int(owner[i]) == rank:
            table[i] = float(10 * i + rank)
    full = parallel.gather_owned_rows(table, owner)
    b, e = parallel.shard_range(7, world, rank)
    local = torch.arange(b, e, dtype=torch.float32).reshape(-1, 1) * torch.ones(1, 2)
    cat = parallel.gather_varlen(local)
    g = torch.ones(4) * (rank + 1)
    work = parallel.allreduce_mlp_grad(g, async_op=True)
    work.wait()
    if rank == 0:
        out["full"] = full.numpy().copy()
        out["cat"] = cat.numpy().copy()
        out["g"] = g.numpy().copy()
    dist.destroy_process_group()


def test_owned_rows_and_varlen_gather_two_ranks():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker_tables, args=(2, port, out), nprocs=2, join=True)
    np.testing.assert_array_equal(out["full"][:, 0], [0, 11, 21, 30, 41])
    np.testing.assert_array_equal(out["cat"][:, 0], np.arange(7))
    np.testing.assert_array_equal(out["g"], [3, 3, 3, 3])
    # without a process group everything is the identity
    t = torch.arange(6.0).reshape(3, 2)
    assert torch.equal(parallel.gather_owned_rows(t, [0, 0, 0]), t) and torch.equal(parallel.gather_varlen(t), t)
    assert parallel.world() == (0, 1) and parallel.allreduce_mlp_grad(t, async_op=True).wait()
