#!/usr/bin/env python3
"""Profiling target for ncu (developer tool): prepares a regime and then issues 10-step launches of the fused rollout
kernel, so that `ncu -k regex:rollout_kernel_multi --launch-skip S --launch-count 1` captures exactly one of them.

    python scripts/prof_target.py settled   # config 2, launches 0..9 = steps 0..99; the robots have settled from launch 6 on
    python scripts/prof_target.py push      # config 2: launches 1..4 (steps 10..49) are in the middle of the pushes
    python scripts/prof_target.py walk      # config 3 (16384 envs, 1.6 s horizon): launches 5..7 = first swing phase
    python scripts/prof_target.py mpc       # dense LIPM-MPC solver, FP64 FMA build then DMMA build (4096 problems, N = 64, 50 repeats)
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linearmpchumanoid_b200 as lmh

mode = sys.argv[1] if len(sys.argv) > 1 else "settled"
if mode == "mpc":
    N, E = 64, 4096
    b = lmh.Batch(2, time_horizon=N * 0.01 * (1 + 1e-9))
    rng = np.random.default_rng(0)
    com, vel = 0.05 * rng.standard_normal((E, 2)), 0.2 * rng.standard_normal((E, 2))
    zx, zy = 0.03 * rng.standard_normal((E, N + 1)), 0.03 * rng.standard_normal((E, N + 1))
    for dm in (False, True):
        print(dm, b.mpc_solve(com, vel, zx, zy, use_dmma=dm, repeats=50)[2], "ms")
    sys.exit(0)
walk = mode == "walk"
E = 16384 if walk else 4096
b = lmh.Batch(E, **({"time_horizon": 1.6} if walk else {}))
if walk:
    b.make_walk_plans(first_env_id=0, seed=0x5EED0003)
else:
    step, dv = lmh.push_plan(range(E))
    b.set_pushes(step, dv)
for k in range(10):
    print(k, b.step_timed(10), "ms")
print(b.stats()["alive"].sum(), "alive")
