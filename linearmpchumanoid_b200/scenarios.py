"""Synthetic scenario generators: thin views of the generators inside liblmh_b200.so
(csrc/lmh_scenarios.hpp, host code behind the C ABI), so that C++ callers, the Python package and the tests all see
the same per-environment integers and doubles.  Everything is keyed by the GLOBAL environment index through a
counter-based generator: a batch gives the same per-environment inputs however it is sharded across GPUs."""
import ctypes as C

import numpy as np

from . import api


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def push_plan(env_ids, seed=0x5EED0002, max_speed=0.25, max_step=50):
    """Config 2: one velocity impulse per env on the base linear x,y velocity; magnitude ~U(0,max_speed),
    direction ~U(0,2pi), at step ~U{0..max_step} (integers, bit-exact across shardings).  -> (step [n] int32, dv [n][6])"""
    L = api.load_library()
    ids = np.ascontiguousarray(np.asarray(list(env_ids) if not isinstance(env_ids, np.ndarray) else env_ids, dtype=np.int64))
    step, dv = np.zeros(ids.size, np.int32), np.zeros((ids.size, 6))
    api._chk(L, L.lmh_push_plan(_p(ids), ids.size, int(seed), float(max_speed), int(max_step), _p(step), _p(dv)))
    return step, dv


def shard_range(n_total, rank, world):
    """contiguous env-index range of a rank."""
    per = n_total // world
    lo = rank * per
    hi = n_total if rank == world - 1 else lo + per
    return lo, hi


def foot_coefficients(cur, des, step_height, T):
    """x(6) y(6) z(8) of one foot over one phase (footRefTrajectory.cpp:4-47)"""
    L = api.load_library()
    cur, des = np.ascontiguousarray(cur, np.float64), np.ascontiguousarray(des, np.float64)
    x, y, z = np.zeros(6), np.zeros(6), np.zeros(8)
    api._chk(L, L.lmh_foot_coeff(_p(cur), _p(des), float(step_height), float(T), _p(x), _p(y), _p(z)))
    return np.concatenate([x, y, z])


def walk_plan(env_ids, seed=0x5EED0003, fixed_len=-1.0, **spec):
    """EXTENSION (BASELINE configs 3-4; upstream declares ZMP::walkZMP but never defines it, zmpGeneration.hpp:15-22): per
    env a plan of n_footsteps steps (right foot first) of length ~U(min_len,max_len), one phase of double support before
    and after.  Keyword arguments are the LmhWalkSpec fields (n_footsteps, time_per_step, dt, horizon_steps, step_height,
    lateral, min_len, max_len, tail_steps).  Returns dict(zmp_x, zmp_y [E][len], phase_T, foot_coef [E][P][2][20],
    contact [E][P] int32 (bit 0 right, bit 1 left), step_len [E])."""
    L = api.load_library()
    ws = api.walk_spec(fixed_len=float(fixed_len), **spec)
    P, Z = np.zeros(1, np.int32), np.zeros(1, np.int32)
    api._chk(L, L.lmh_walk_plan_dims(C.byref(ws), _p(P), _p(Z)))
    ids = np.ascontiguousarray(np.asarray(list(env_ids) if not isinstance(env_ids, np.ndarray) else env_ids, dtype=np.int64))
    E, P, Z = ids.size, int(P[0]), int(Z[0])
    zx, zy = np.zeros((E, Z)), np.zeros((E, Z))
    fc, ct, lens = np.zeros((E, P, 2, 20)), np.zeros((E, P), np.int32), np.zeros(E)
    api._chk(L, L.lmh_walk_plan(C.byref(ws), _p(ids), E, int(seed), _p(zx), _p(zy), _p(fc), _p(ct), _p(lens)))
    return dict(zmp_x=zx, zmp_y=zy, phase_T=ws.time_per_step, foot_coef=fc, contact=ct, step_len=lens)
