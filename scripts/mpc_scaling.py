#!/usr/bin/env python3
"""Developer tool (GPU box: PYTHONPATH=. python scripts/mpc_scaling.py): dense LIPM-MPC solver, FP64 FMA build against DMMA
build, over the number of problems -- the FMA build maps one lane to a right-hand-side column (32 columns per warp), the
DMMA build one warp to 8 columns, so at small batches the FMA build is also short of warps; the large-batch rows are the
fair comparison of the two pipes."""
import json, sys
import numpy as np
import linearmpchumanoid_b200 as lmh
out = []
for N in (16, 64):
    b = lmh.Batch(2, time_horizon=N * 0.01 * (1 + 1e-9))
    fp64, _, _ = lmh.measure_fp64_peak(0)
    for E in (4096, 32768, 262144):
        rng = np.random.default_rng(0)
        com, vel = 0.05 * rng.standard_normal((E, 2)), 0.2 * rng.standard_normal((E, 2))
        zx, zy = 0.03 * rng.standard_normal((E, N + 1)), 0.03 * rng.standard_normal((E, N + 1))
        row = {"N": N, "problems": E}
        for name, dm in (("dfma", False), ("dmma", True)):
            R = 20
            b.mpc_solve(com, vel, zx, zy, use_dmma=dm, repeats=2)
            t = min(b.mpc_solve(com, vel, zx, zy, use_dmma=dm, repeats=R)[2] for _ in range(3))
            t1 = min(b.mpc_solve(com, vel, zx, zy, use_dmma=dm, repeats=1)[2] for _ in range(3))
            rate = E * (R - 1) / ((t - t1) * 1e-3)
            row[name + "_qps_per_s"] = rate
            row[name + "_tflops"] = rate * (8.0 * (N + 1) ** 2 + 8.0 * (N + 1)) / 1e12
        row["dmma_over_dfma"] = row["dmma_qps_per_s"] / row["dfma_qps_per_s"]
        row["fp64_peak_tflops"] = fp64
        print(json.dumps(row), flush=True)
