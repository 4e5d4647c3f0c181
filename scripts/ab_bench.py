#!/usr/bin/env python3
"""Developer tool: A/B timing of kernel variants.  Builds liblmh_b200 with extra -D flags into
linearmpchumanoid_b200/variants/ (build, run here: nvcc cross-compiles) and times the config-2 rollout of each
variant on the GPU box (run), one process per variant.  Not part of the product or of bench.py.

    python scripts/ab_bench.py build  tag1:-DLMH_X=0 tag2:-DLMH_X=1,-DLMH_Y=2 ...
    python scripts/ab_bench.py run    [--envs 4096] [--rollout 200] [--reps 5] [--workload push|walk]
"""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VAR = os.path.join(ROOT, "linearmpchumanoid_b200", "variants")
sys.path.insert(0, ROOT)


def build(specs):
    from linearmpchumanoid_b200 import api
    os.makedirs(VAR, exist_ok=True)
    src = os.path.join(ROOT, "linearmpchumanoid_b200", "csrc", "lmh_capi.cu")
    procs = []
    for spec in specs:
        tag, _, flags = spec.partition(":")
        out = os.path.join(VAR, f"liblmh_b200_{tag}.so")
        cmd = ["nvcc"] + api.NVCC_FLAGS + ["-Xptxas=-v", "-I", os.path.join(ROOT, "include"), "-o", out, src] + [f for f in flags.split(",") if f]
        procs.append((tag, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for tag, p in procs:
        log = p.communicate()[0]
        if p.returncode != 0:
            raise SystemExit(f"{tag}: nvcc failed\n{log}")
        lines = log.split("\n")
        for i, l in enumerate(lines):
            if "rollout_kernel_multi" in l and "Function properties" in l:
                print(tag, "|", lines[i + 1].strip(), "|", lines[i + 2].strip())


def one(path, envs, rollout, reps, workload):
    import numpy as np
    from linearmpchumanoid_b200 import api
    import linearmpchumanoid_b200 as lmh
    api.LIB_PATH = path
    api._lib = None
    walk = workload == "walk"
    b = lmh.Batch(envs, **({"time_horizon": 1.6} if walk else {}))
    if walk:
        wp = lmh.walk_plan(range(envs))
        b.set_walk_plans(wp["zmp_x"], wp["zmp_y"], wp["phase_T"], wp["foot_coef"], wp["contact"], plan_ids=np.arange(envs))
    else:
        step, dv = lmh.push_plan(range(envs))
        b.set_pushes(step, dv)
    ms = []
    for r in range(reps + 1):
        b.reset()
        b.sync()
        t = b.step_timed(rollout)
        if r:
            ms.append(t)
    settled = min(b.step_timed(50) for _ in range(3))
    st = b.stats()
    x = b.get_state()[0]
    print(json.dumps({"variant": os.path.basename(path), "workload": workload, "env_steps_per_s": envs * rollout / (min(ms) * 1e-3),
                      "ms": min(ms), "settled_env_steps_per_s": envs * 50 / (settled * 1e-3),
                      "checksum": float(np.abs(x).sum()), "alive": int(st["alive"].sum()), "qp_fail": int(st["qp_fail"].sum())}), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "one":
        one(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6])
    else:
        a = sys.argv[2:]
        opt = lambda k, d: a[a.index(k) + 1] if k in a else d
        for f in sorted(os.listdir(VAR)):
            if f.endswith(".so"):
                subprocess.run([sys.executable, __file__, "one", os.path.join(VAR, f), opt("--envs", "4096"), opt("--rollout", "200"),
                                opt("--reps", "5"), opt("--workload", "push")])
