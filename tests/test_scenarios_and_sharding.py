"""Scenario generators and the N > 1 host logic (contiguous env shards, final gather) on CPU with gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import linearmpchumanoid_b200 as lmh


def test_push_plan_is_keyed_by_global_env_id():
    step, dv = lmh.push_plan(range(64))
    assert step.dtype == np.int32 and step.min() >= 0 and step.max() <= 50
    mag = np.hypot(dv[:, 0], dv[:, 1])
    assert mag.max() <= 0.25 and (dv[:, 2:] == 0).all()
    for world in (2, 4, 8):                          # any sharding reproduces the same per-env integers and doubles
        parts = [lmh.push_plan(range(*lmh.shard_range(64, r, world))) for r in range(world)]
        assert np.array_equal(np.concatenate([p[0] for p in parts]), step)
        assert np.array_equal(np.concatenate([p[1] for p in parts]), dv)
    assert not np.array_equal(lmh.push_plan(range(64), seed=1)[0], step)


def test_shard_ranges_cover_everything_once():
    for n, w in ((4096, 1), (4096, 8), (1000, 3), (7, 8)):
        seen = np.zeros(n, int)
        for r in range(w):
            lo, hi = lmh.shard_range(n, r, w)
            seen[lo:hi] += 1
        assert (seen == 1).all()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_total, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    lo, hi = lmh.shard_range(n_total, rank, world)
    step, dv = lmh.push_plan(range(lo, hi))
    # stand-in for the per-env statistics a rank's GPU would produce: a pure function of the global env id
    local = np.stack([np.arange(lo, hi, dtype=np.float64), step.astype(np.float64), dv[:, 0], dv[:, 1]], axis=1)
    gathered = bench.gather_stats(torch.from_numpy(local), rank, world, "cpu")
    ms = bench.max_over_ranks(10.0 + rank, "cpu")
    if rank == 0:
        np.save(out_path, np.concatenate([gathered.numpy().ravel(), [ms]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_with_gloo(tmp_path):
    out = str(tmp_path / "g.npy")
    mp.spawn(_worker, args=(2, _free_port(), 100, out), nprocs=2, join=True)
    flat = np.load(out)
    g, ms = flat[:-1].reshape(100, 4), flat[-1]
    step, dv = lmh.push_plan(range(100))
    assert np.array_equal(g[:, 0], np.arange(100.0)) and np.array_equal(g[:, 1], step) and np.array_equal(g[:, 2:], dv[:, :2])
    assert ms == 11.0                                  # max over ranks
