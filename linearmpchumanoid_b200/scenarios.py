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
