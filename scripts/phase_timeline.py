#!/usr/bin/env python3
"""Developer tool: where one stage evaluation spends its time.  Needs the debug build with clock64 stamps at the phase
boundaries (python scripts/ab_bench.py build tl:-DLMH_TIMELINE=1; the stamps travel in the tau_k1 field of the step log),
run on the GPU box: python scripts/phase_timeline.py"""
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
from linearmpchumanoid_b200 import api
import linearmpchumanoid_b200 as lmh
api.LIB_PATH = os.path.join(os.getcwd(), "linearmpchumanoid_b200/variants/liblmh_b200_tl.so"); api._lib = None
names = ["stage start", "fk+twist done", "handover2 passed", "rnea done", "handover3 passed (crba in throughput mode)", "centroidal done", "mpc_pd done", "assemble_ga done", "factor done", "qp solved (before backsolve)", "backsolve done", "output done"]
for mode, kw in (("latency", dict(envs_per_block=-1)), ("throughput-1env", {}), ("throughput-1480", {})):
    E = 1480 if mode.endswith("1480") else 1
    b = lmh.Batch(E, log_steps=60, **kw)
    b.step(60); b.sync()
    lg = b.log(0, 60)
    st = lg["tau_k1"][50:60, 0, :16]          # env 0, steps 50..59 (stamps of stage 3 of each step)
    d = np.diff(st[:, :12], axis=1).mean(0) / 1965.0    # us
    print(mode, "stage total %.2f us; step (12->13) %.2f us" % ((st[:, 11] - st[:, 0]).mean() / 1965.0, (st[:, 13] - st[:, 12]).mean() / 1965.0))
    for n, x in zip(names[1:], d): print("   %-45s %.2f us" % (n, x))
