"""Synthetic scenario generators.  Everything is keyed by the GLOBAL environment
index through a counter-based generator, so a batch gives the same per-environment inputs however it
is sharded across GPUs."""
import numpy as np

_M64 = (1 << 64) - 1


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & _M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def _u01(seed, env, stream):
    """uniform [0,1) from (seed, env id, stream id): 53 high bits of a splitmix64 hash."""
    h = _splitmix64((seed * 0x100000001B3 + env * 0x9E3779B1 + stream * 0x85EBCA77) & _M64)
    h = _splitmix64(h)
    return (h >> 11) * (1.0 / (1 << 53))


def push_plan(env_ids, seed=0x5EED0002, max_speed=0.25, max_step=50):
    """Config 2: one velocity impulse per env on the base linear x,y velocity; magnitude ~U(0,max_speed),
    direction ~U(0,2pi), at step ~U{0..max_step} (integers, bit-exact across shardings)."""
    env_ids = np.asarray(env_ids, dtype=np.int64)
    step = np.zeros(env_ids.size, np.int32)
    dv = np.zeros((env_ids.size, 6))
    for i, e in enumerate(env_ids.tolist()):
        mag = max_speed * _u01(seed, e, 0)
        ang = 2.0 * np.pi * _u01(seed, e, 1)
        step[i] = int(_u01(seed, e, 2) * (max_step + 1))
        dv[i, 0] = mag * np.cos(ang)
        dv[i, 1] = mag * np.sin(ang)
    return step, dv


def shard_range(n_total, rank, world):
    """contiguous env-index range of a rank."""
    per = n_total // world
    lo = rank * per
    hi = n_total if rank == world - 1 else lo + per
    return lo, hi


# ---------------------------------------------------------------------------------------------------------------
# EXTENSION: walking plans (BASELINE configs 3-4).  The reference declares ZMP::walkZMP but never defines it
# (zmpGeneration.hpp:15-22) and has no contact switching; this is our definition.
def _poly_fit(pt, pv, vt, vv, at, av):
    """boundary-value polynomial, ascending powers (generalizedFunctions.cpp:103-163)"""
    n = len(pt) + len(vt) + len(at)
    A, b = np.zeros((n, n)), np.zeros(n)
    row = 0
    for t, v in zip(pt, pv):
        A[row] = [t ** j for j in range(n)]; b[row] = v; row += 1
    for t, v in zip(vt, vv):
        A[row] = [0.0] + [j * t ** (j - 1) for j in range(1, n)]; b[row] = v; row += 1
    for t, v in zip(at, av):
        A[row] = [0.0, 0.0] + [j * (j - 1) * t ** (j - 2) for j in range(2, n)]; b[row] = v; row += 1
    return np.linalg.solve(A, b)


def foot_coefficients(cur, des, step_height, T):
    """x(6) y(6) z(8) of one foot over one phase (footRefTrajectory.cpp:4-47)"""
    t2, z2, t3, z3 = [0.0, T], [0.0, 0.0], [0.0, T / 2, T], [0.0, 0.0, 0.0]
    x = _poly_fit(t2, [cur[0], des[0]], t2, z2, t2, z2)
    y = _poly_fit(t2, [cur[1], des[1]], t2, z2, t2, z2)
    z = _poly_fit(t3, [cur[2], step_height, des[2]], t3, z3, t2, z2)
    return np.concatenate([x, y, z])


def walk_plan(env_ids, seed=0x5EED0003, n_footsteps=4, time_per_step=0.4, dt=0.01, horizon_steps=160,
              step_height=0.02, lateral=0.05, min_len=0.02, max_len=0.05, tail_steps=40):
    """Config 3: per env a plan of n_footsteps steps (right foot first) of length ~U(min_len,max_len), one phase of
    double support before and after.  Returns dict(zmp_x, zmp_y [E][len], phase_T, foot_coef [E][P][2][20],
    contact [E][P] int32 (bit 0 right, bit 1 left), step_len [E])."""
    env_ids = np.asarray(env_ids, dtype=np.int64)
    P = n_footsteps + 2
    per = int(round(time_per_step / dt))
    zlen = P * per + tail_steps + horizon_steps + 2
    E = env_ids.size
    zx, zy = np.zeros((E, zlen)), np.zeros((E, zlen))
    fc, ct = np.zeros((E, P, 2, 20)), np.zeros((E, P), np.int32)
    lens = np.zeros(E)
    for i, e in enumerate(env_ids.tolist()):
        L = min_len + (max_len - min_len) * _u01(seed, e, 0)
        lens[i] = L
        foot = [np.array([0.0, -lateral, 0.0]), np.array([0.0, lateral, 0.0])]     # right, left
        for p in range(P):
            swing = -1 if p == 0 or p == P - 1 else (p - 1) % 2                      # right foot swings first
            ct[i, p] = 3 if swing < 0 else (2 if swing == 0 else 1)
            target = [f.copy() for f in foot]
            if swing >= 0:
                target[swing][0] = foot[1 - swing][0] + L
            for s in range(2):
                fc[i, p, s] = foot_coefficients(foot[s], target[s], step_height if s == swing else 0.0, time_per_step)
            if swing < 0:
                z = 0.5 * (foot[0][:2] + foot[1][:2])
            else:
                z = foot[1 - swing][:2]
            k0, k1 = p * per, (p + 1) * per if p < P - 1 else zlen
            zx[i, k0:k1], zy[i, k0:k1] = z[0], z[1]
            foot = target
    return dict(zmp_x=zx, zmp_y=zy, phase_T=time_per_step, foot_coef=fc, contact=ct, step_len=lens)
