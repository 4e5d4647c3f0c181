#!/usr/bin/env python3
"""Developer tool (GPU box: PYTHONPATH=. python scripts/latency_mode_crossover.py): throughput kernel against the opt-in
latency kernel (envs_per_block = -1) over the batch size -- where the helper warps stop paying."""
import json
import linearmpchumanoid_b200 as lmh
for E in (1, 148, 296, 592, 888, 1184, 1480, 2960):
    row = {"envs": E}
    for name, kw in (("throughput_kernel", {}), ("latency_kernel", {"envs_per_block": -1})):
        b = lmh.Batch(E, **kw)
        b.step(20)
        ms = min(b.step_timed(100) for _ in range(3))
        row[name + "_us_per_step"] = ms * 10.0
        row[name + "_env_steps_per_s"] = E * 100 / (ms * 1e-3)
        b.close()
    print(json.dumps(row), flush=True)
