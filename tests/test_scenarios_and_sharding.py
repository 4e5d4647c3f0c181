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


def test_walk_plan_is_keyed_by_global_env_id_and_matches_its_definition():
    """the walking-plan generator behind the C ABI (ZMP::walkZMP + per-phase foot polynomials + contact masks)"""
    wp = lmh.walk_plan(range(40), seed=0x5EED0003)
    P, per = 6, 40
    assert wp["contact"].shape == (40, P) and (wp["contact"] == [3, 2, 1, 2, 1, 3]).all()      # right foot swings first
    assert wp["zmp_x"].shape[1] == P * per + 40 + 160 + 2 and wp["phase_T"] == 0.4
    assert (wp["step_len"] >= 0.02).all() and (wp["step_len"] <= 0.05).all() and len(np.unique(wp["step_len"])) == 40
    for world in (2, 4, 8):                          # any sharding reproduces the same per-env integers and doubles
        parts = [lmh.walk_plan(range(*lmh.shard_range(40, r, world)), seed=0x5EED0003) for r in range(world)]
        for k in ("zmp_x", "zmp_y", "foot_coef", "contact", "step_len"):
            assert np.array_equal(np.concatenate([p[k] for p in parts]), wp[k]), k
    # definition, environment 5: ZMP under the stance foot in single support, between the feet in double support;
    # swing polynomials start at the foot's place, end one step ahead of the other foot, peak at step_height
    L = wp["step_len"][5]
    zx, zy, fc = wp["zmp_x"][5], wp["zmp_y"][5], wp["foot_coef"][5]
    assert (zx[:per] == 0).all() and (zy[:per] == 0).all()                       # double support at the start
    assert (zy[per:2 * per] == 0.05).all() and (zx[per:2 * per] == 0).all()      # right foot swings: ZMP under the left foot
    assert (zy[2 * per:3 * per] == -0.05).all() and np.allclose(zx[2 * per:3 * per], L)
    pv = lambda c, t: np.polyval(c[::-1], t)
    rx, rz = fc[1, 0, 0:6], fc[1, 0, 12:20]                                      # phase 1, right foot
    assert abs(pv(rx, 0.0)) < 1e-14 and abs(pv(rx, 0.4) - L) < 1e-12 and abs(pv(rz, 0.2) - 0.02) < 1e-12 and abs(pv(rz, 0.4)) < 1e-12
    assert np.abs(fc[1, 1, 0:6] - [0, 0, 0, 0, 0, 0]).max() == 0 and fc[1, 1, 6] == 0.05        # the left foot stays put
    fixed = lmh.walk_plan([3, 9], fixed_len=0.03)
    assert (fixed["step_len"] == 0.03).all() and np.array_equal(fixed["zmp_x"][0], fixed["zmp_x"][1])


def test_foot_coefficients_against_numpy():
    """lmh_foot_coeff / lmh_poly_fit (findPolyCoeff, generalizedFunctions.cpp:103-163) against numpy's solve of the same
    boundary-value system, and against the reference's stand-task coefficients"""
    golden = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz"))
    c = np.concatenate([lmh.foot_coefficients([0, -0.05, 0], [0, -0.05, 0], 0.0, 5.0), lmh.foot_coefficients([0, 0.05, 0], [0, 0.05, 0], 0.0, 5.0)])
    assert np.abs(c - golden["foot_coef"]).max() < 1e-12
    T, cur, des, hgt = 0.4, [0.01, -0.05, 0.0], [0.05, -0.04, 0.0], 0.02
    got = lmh.foot_coefficients(cur, des, hgt, T)

    def fit(pt, pv, vt, vv, at, av):
        n = len(pt) + len(vt) + len(at)
        A = np.array([[t ** j for j in range(n)] for t in pt] + [[j * t ** (j - 1) if j else 0.0 for j in range(n)] for t in vt]
                     + [[j * (j - 1) * t ** (j - 2) if j > 1 else 0.0 for j in range(n)] for t in at])
        return np.linalg.solve(A, np.array(list(pv) + list(vv) + list(av), float))
    x = fit([0, T], [cur[0], des[0]], [0, T], [0, 0], [0, T], [0, 0])
    z = fit([0, T / 2, T], [cur[2], hgt, des[2]], [0, T / 2, T], [0, 0, 0], [0, T], [0, 0])
    assert np.abs(got[:6] - x).max() < 1e-9 * np.abs(x).max() and np.abs(got[12:] - z).max() < 1e-9 * np.abs(z).max()
