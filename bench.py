#!/usr/bin/env python3
"""bench.py -- closed-loop env-steps/s of the batched NAO controller-and-simulator (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs E] [--rollout S]

Workload (config.workload), default: BASELINE.json configs[1] = NAO balance under random push perturbations, E = 4096
environments per GPU, one push per env (seed 0x5EED0002, keyed by the global env id), S = 200
RK4 steps per rollout.  One bench "step" = one whole rollout of the batch: restore the initial state, run S
fused RK4 steps per env (4 stage evaluations each: kinematics + dynamics + LIPM-MPC + WBC QP + inverse
dynamics), read the statistics back.  --workload walk = configs[2] (16384 envs x 240 steps, 1.6 s horizon, seed
0x5EED0003), sweep = configs[3] (32768 envs per GPU x 240 steps, walking + pushes, seed 0x5EED0004), qp = configs[4]
(cold WBC QPs and dense LIPM-MPC solves on frozen states); --envs / --rollout override the sizes.

  value : env-steps/s, whole job over all N GPUs, initial state resident in HBM (device-to-device restore),
          timed with CUDA events on the launching stream, max over ranks.
  e2e   : the same through the C ABI with HOST buffers: pinned host -> device copy of the states and pushes,
          the rollout, device -> host read of final states and statistics, all inside the timed region.
  roofline : FP64-FMA bound (0.51 MFLOP and 2-5 KB of HBM traffic per env-step); peak measured
          in this run by an FP64 FMA micro-kernel (MEASURED_PEAKS.json has no FP64 entry); HBM fraction beside it.
  cpu_baseline : the oracle port of the reference's CPU path on the host cores, bounded sample (rank 0, N = 1).

--impl reference times the reference's own CPU implementation (oracle/_ref: upstream sources compiled against
shim Eigen/qpOASES headers; falls back to the oracle port) on the host cores, same metric and config.
Multi-GPU: one process per GPU under torchrun; environments shard by index, no per-step communication; the only
collective is the final gather of per-env statistics (NCCL all_gather), reported separately as gather_ms.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {"qp": "QP microbenchmark on frozen states (BASELINE configs[4])",
             "push": "NAO balance under random push perturbations (BASELINE configs[1])",
             "walk": "NAO LIPM-MPC walking, random footstep plans, 1.6 s horizon (BASELINE configs[2]; extension: upstream has no walking)",
             "sweep": "NAO walking robustness Monte Carlo: footstep plans + pushes (BASELINE configs[3]; extension)"}
# per workload: environments per GPU, RK4 steps per rollout, scenario seed, MPC horizon in samples
DEFAULTS = {"push": (4096, 200, 0x5EED0002, 50), "walk": (16384, 240, 0x5EED0003, 160), "sweep": (32768, 240, 0x5EED0004, 160),
            "qp": (4096, 50, 0x5EED0002, 50)}
FLOP_PER_ENV_STEP = 0.51e6          # SURVEY.md 8(d) / App. C, active-set changes k = 0 (conservative)
STATE_BYTES_PER_ENV = (60 + 30 + 1) * 8 * 2 + 48 + 24   # state read+write, stats, CoM per launch
# dram__bytes_read.sum + dram__bytes_write.sum of rollout_kernel_multi from the committed ncu --set full captures (the
# traffic of a launch does not depend on the number of steps it fuses: the state stays on chip).  Walking reads each
# environment's own plan (ZMP reference + per-phase foot polynomials) as well.
NCU_DRAM_BYTES_PER_LAUNCH = {"push": 3572224 + 0, "walk": 68396544 + 2678272, "sweep": None}
NCU_TRAFFIC_SOURCE = {"push": "profiles/r02_rollout_push.txt: dram read 3.57 MB + write 0 for a 10-step launch of 4096 envs "
                              "(algorithmic: 6.3 MB of state and statistics per launch, independent of the step count; the "
                              "written state was still in the 126 MB L2 when the kernel ended)",
                      "walk": "profiles/r02_rollout_walk.txt: dram read 68.4 MB + write 2.7 MB for a 10-step launch of 16384 envs "
                              "(algorithmic: 25 MB of state and statistics + 16384 plans x 9 KB = 147 MB if every plan were read whole; "
                              "a 10-step launch touches only the current ZMP window and phase)",
                      "sweep": "no capture of the 32768-environment sweep launch (same kernel and plan layout as the walking workload)"}


# ---------------------------------------------------------------- distributed helpers (also unit-tested with gloo)
def gather_stats(local, rank, world, device):
    """final gather of per-env statistics [n_local, k] -> [n_total, k] on every rank (the path's only collective)"""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    local = local.to(device).contiguous()
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local)
    return torch.cat(parts, 0).cpu()


def max_over_ranks(ms, device):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms)
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _barrier(device):
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if device != "cpu":
        torch.cuda.synchronize()


# ---------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc, self.path = None, f"/tmp/lmh_clocks_{os.getpid()}.csv"
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons, power = [], [], set(), []
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[0])); smax.append(float(p[1])); power.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [c for c, w in zip(sm, power) if w > 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


# ---------------------------------------------------------------- CPU arms
def cpu_rollout_rate(lib, kind, n_env, n_steps, threads, first_env=0):
    """env-steps/s of a CPU implementation over the first n_env environments of the workload"""
    import linearmpchumanoid_b200 as lmh
    kw = dict(second_wbc=0, mpc_rebuild_hessian=0) if kind == "port" else {}
    cfg = lib.default_config(**kw)
    step, dv = lmh.push_plan(range(first_env, first_env + n_env))
    t0 = time.perf_counter()
    x, com, st, fails = lib.batch_rollout(cfg, n_env, n_steps, push_step=step, push_dv=dv, n_threads=threads)
    dt = time.perf_counter() - t0
    return n_env * n_steps / dt, dt, x


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc_py
    lib, kind, note = None, "reference", ""
    try:
        lib = orc_py.reference()
    except Exception:
        lib = None
    if lib is not None:
        note = ("reference sources + NAIVE shims: upstream src/*.cpp + apps/offline/main.cpp compiled unmodified (-O2) against "
                "this repository's stand-ins for Eigen (eager, heap-allocating) and qpOASES (oracle/refbuild; the real "
                "libraries are absent and would be faster); solves each stage's WBC QP twice and rebuilds the MPC Hessian "
                "per call, as upstream does.  The tighter CPU baseline is the oracle port (cpu_baseline of the GPU arm).")
    else:
        orc_py.build()
        lib, kind = orc_py.oracle(), "port"
        note = "oracle/_ref unavailable on this box: oracle port of the reference CPU path (one WBC solve per stage)"
    cores = os.cpu_count() or 1
    n_env = max(cores, 8)
    S = max(4, min(args.rollout, 40))
    rates = []
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_rollout_rate(lib, kind, n_env, 2, cores)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        r, dt, _ = cpu_rollout_rate(lib, kind, n_env, S, cores)
        rates.append(r)
        if time.perf_counter() - t_all > 150:
            break
    value = float(np.mean(rates))
    line = {"impl": "reference", "metric": "closed-loop env-steps/sec (NAO, FP64)", "value": value, "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": len(rates), "warmup": args.warmup, "ms_per_step": 1e3 * n_env * S / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "NAO balance under random push perturbations (BASELINE configs[1])",
                       "envs_per_step_sample": n_env, "rk4_steps_per_step_sample": S, "seed": "0x5EED0002"},
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": kind,
                             "sample": f"first {n_env} envs x {S} RK4 steps per bench step, {cores} host threads", "note": note},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)
    return 0


# ---------------------------------------------------------------- config 5: QP micro-benchmark
def run_qp_microbench(args):
    """BASELINE configs[4]: frozen states taken at step 50 of config 2; cold-started WBC QPs (74 var / 18 eq / 32 ineq in
    the reference's form) and LIPM-MPC evaluations on the GPU vs the oracle's solvers on the host cores."""
    import torch
    import linearmpchumanoid_b200 as lmh
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    E, R = args.envs, 50
    b = lmh.Batch(E, threads_per_env=args.threads_per_env)
    step, dv = lmh.push_plan(range(E))
    b.set_pushes(step, dv)
    b.step(50)                                      # frozen states = config 2 at step 50
    X, VH, T = b.get_state()
    b.qp_bench(0, 2)
    ms_qp, a, f = b.qp_bench(0, R)
    ms_model, _, _ = b.qp_bench(0, 1)
    ms_mpc, _, _ = b.qp_bench(1, 4 * R)
    qps = E * (R - 1) / ((ms_qp - ms_model) * 1e-3)
    mpcs = E * 4 * R / ((ms_mpc - ms_model) * 1e-3) if ms_mpc > ms_model else float("nan")
    # oracle on the host: the same QPs through its own solver, bounded sample
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc_py
    orc_py.build()
    O = orc_py.oracle()
    so = O.sim()
    n = min(E, 200)
    t0, worst = time.perf_counter(), 0.0
    for e in range(n):
        so.set_state(X[e], float(T[e]), VH[e])
        d = so.stage(X[e], float(T[e]))
        xa, xf = d["x"][:30], d["x"][30:42]
        worst = max(worst, float(np.abs(a[e] - xa).max() / max(1.0, np.abs(xa).max())),
                    float(np.abs(f[e] - xf).max() / max(1.0, np.abs(xf).max())))
    cpu_stage_rate = n / (time.perf_counter() - t0)
    # condensed LIPM-MPC over the horizons BASELINE names (16-64 samples) on the same frozen states: evaluations/s and the
    # MPC outputs [pos+, vel+, u0] of both axes against the oracle's Mpc3dLip (dense Hessian solve) at that horizon
    mpc_sweep, RM = {}, 200
    fp64_tf, _, _ = lmh.measure_fp64_peak(0)
    for N in (16, 32, 64):
        hz = N * b.cfg.dt * (1.0 + 1e-9)            # (int)(horizon / dt) == N, as mpcLinearPendulum.hpp:49 computes it
        bn = lmh.Batch(E, threads_per_env=args.threads_per_env, time_horizon=hz)
        bn.set_state(X, VH, T)
        ms_m0, _, _ = bn.qp_bench(0, 1)
        ms_m1, _, _ = bn.qp_bench(1, 4 * R)
        dn = bn.stage()
        son = O.sim(time_horizon=hz)
        err = 0.0
        for e in range(min(E, 50)):
            son.set_state(X[e], float(T[e]), VH[e])
            do = son.stage(X[e], float(T[e]))
            err = max(err, float(np.abs(dn["xRef"][e] - do["xRef"]).max()), float(np.abs(dn["yRef"][e] - do["yRef"]).max()))
        # the dense solver on the device (lmh_mpc_solve: gradient by a dense product, shared Cholesky factor, forward and
        # back substitution per problem), FP64 FMA build and FP64 tensor-core (DMMA m8n8k4) build of the same algorithm
        com, vel = np.ascontiguousarray(dn["CoM"][:, :2]), np.ascontiguousarray(dn["comVel"][:, :2])
        zwin = np.zeros((E, N + 1))
        dense = {}
        for name, dm in (("dfma", False), ("dmma", True)):
            bn.mpc_solve(com, vel, zwin, zwin, use_dmma=dm, repeats=2)
            xr, yr, ms_d = bn.mpc_solve(com, vel, zwin, zwin, use_dmma=dm, repeats=RM)
            _, _, ms_1 = bn.mpc_solve(com, vel, zwin, zwin, use_dmma=dm, repeats=1)
            rate = E * (RM - 1) / ((ms_d - ms_1) * 1e-3)
            flop = 8.0 * (N + 1) ** 2 + 8.0 * (N + 1)            # two axes, shared factor
            dense[name] = {"qps_per_s": rate, "tflops": rate * flop / 1e12, "frac_of_fp64_peak": rate * flop / 1e12 / fp64_tf,
                           "max_abs_err_vs_gain_form": float(max(np.abs(xr - dn["xRef"]).max(), np.abs(yr - dn["yRef"]).max()))}
        mpc_sweep[str(N)] = {"gain_form_evaluations_per_s": E * 4 * R / ((ms_m1 - ms_m0) * 1e-3) if ms_m1 > ms_m0 else float("nan"),
                             "dense_solve": dense, "max_abs_err_refs_vs_oracle": err}
        bn.close()
    line = {"metric": "WBC QP solves/sec (74 var, 18 eq, 32 ineq; cold start, frozen states)", "value": qps, "unit": "QPs/s",
            "n_gpus": 1, "steps": R, "warmup": 2, "ms_per_step": (ms_qp - ms_model) / (R - 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "QP microbenchmark on frozen states (BASELINE configs[4])", "envs": E,
                       "states": "config 2 at step 50", "repeats": R},
            "mpc_gain_form_evaluations_per_s": mpcs, "mpc_horizon_sweep": mpc_sweep, "fp64_peak_tflops_builder_measured": fp64_tf, "max_rel_err_a_f_vs_oracle": worst, "oracle_sample_envs": n,
            "cpu_baseline": {"value": cpu_stage_rate, "unit": "stage evaluations/s (model + MPC + QP)", "cores": 1, "kind": "port",
                             "sample": f"{n} frozen states, oracle stage evaluation (its QP is ~85% of a stage)"},
            "gpu_launches": 4 + 3 * (3 + 6)}
    _emit(line)
    b.close()
    return 0


# ---------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import linearmpchumanoid_b200 as lmh

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))

    E, S, K, W = args.envs, args.rollout, args.steps, max(args.warmup, 0)
    seed = DEFAULTS[args.workload][2]
    lo = rank * E                                             # weak scaling: E envs per GPU, global ids lo..lo+E
    walk = args.workload in ("walk", "sweep")
    b = lmh.Batch(E, device=local_rank, threads_per_env=args.threads_per_env, **({"time_horizon": 1.6} if walk else {}))
    q0, zcom, mass, _ = b.nominal()
    if walk:      # configs 3-4 (EXTENSION: the reference has no walking): per-env footstep plans, 1.6 s MPC horizon,
        b.make_walk_plans(first_env_id=lo, seed=seed)          # generated inside the library, keyed by the global env id
    if args.workload in ("push", "sweep"):
        push_step, push_dv = lmh.push_plan(range(lo, lo + E), seed=seed)
    else:
        push_step, push_dv = np.full(E, -1, np.int32), np.zeros((E, 6))
    b.set_pushes(push_step, push_dv)

    # initial state: pinned host copy (e2e leg) and device-resident template (value leg)
    x0_host = torch.zeros(E, 60, dtype=torch.float64).pin_memory()
    x0_host[:, :30] = torch.from_numpy(q0)
    vh0_host = torch.zeros(E, 30, dtype=torch.float64).pin_memory()
    t0_host = torch.zeros(E, dtype=torch.float64).pin_memory()
    x0_dev, vh0_dev, t0_dev = x0_host.to(dev), vh0_host.to(dev), t0_host.to(dev)
    x_out = torch.zeros(E, 60, dtype=torch.float64).pin_memory()
    stats_out = np.zeros(E, lmh.STATS_DTYPE)
    stats_pinned = torch.zeros(stats_out.nbytes, dtype=torch.uint8).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    torch.cuda.synchronize()

    def rollout_resident():
        b.set_state_dev(x0_dev.data_ptr(), vh0_dev.data_ptr(), t0_dev.data_ptr())
        b.reset_stats()
        b.step(S, sync=False)

    def rollout_e2e():
        b.set_state_ptr(x0_host.data_ptr(), vh0_host.data_ptr(), t0_host.data_ptr())
        b.set_pushes(push_step, push_dv)
        b.reset_stats()
        b.step(S, sync=False)
        b.get_state_ptr(x_out.data_ptr())
        b.get_stats_ptr(stats_pinned.data_ptr())

    def timed(fn, n):
        total = 0.0
        for _ in range(n):
            flush.zero_()
            torch.cuda.synchronize()
            b.timer_start()
            fn()
            total += b.timer_stop()
        return total

    for _ in range(max(W, 3)):
        rollout_resident()
    b.sync()
    launches0 = b.launch_info()["launches"]
    sampler = ClockSampler(local_rank) if rank == 0 else None
    _barrier(dev)
    wall0 = time.perf_counter()
    ms_total = timed(rollout_resident, K)
    b.sync()
    _barrier(dev)
    wall_ms = 1e3 * (time.perf_counter() - wall0)
    launches = b.launch_info()["launches"] - launches0
    ms_total = max_over_ranks(ms_total, dev)

    # kernel-only duration of the dominant kernel (the fused rollout), live, CUDA events on its stream
    kern_ms = []
    for _ in range(min(K, 3)):
        b.set_state_dev(x0_dev.data_ptr(), vh0_dev.data_ptr(), t0_dev.data_ptr())
        b.reset_stats()
        b.sync()
        kern_ms.append(b.step_timed(S))
    kern = float(np.mean(kern_ms))

    # end-to-end leg through host buffers
    for _ in range(2):
        rollout_e2e()
    _barrier(dev)
    e2e_ms = max_over_ranks(timed(rollout_e2e, K), dev)
    clocks = sampler.stop() if sampler else None
    stats_out = np.frombuffer(stats_pinned.numpy().tobytes(), dtype=lmh.STATS_DTYPE)
    h2d = E * (60 + 30 + 1) * 8 + E * (4 + 48)
    d2h = E * 60 * 8 + stats_out.nbytes

    # final gather of the per-env statistics: the only collective of the path
    local_stats = torch.tensor(np.stack([stats_out["max_abs_tau"], stats_out["max_com_err"], stats_out["max_abs_qpp"],
                                         stats_out["qp_iters"].astype(np.float64), stats_out["qp_fail"].astype(np.float64),
                                         stats_out["steps"].astype(np.float64), stats_out["alive"].astype(np.float64)], 1))
    gather_stats(local_stats, rank, world, dev)          # untimed: the first collective of a process sets NCCL up
    _barrier(dev)
    tg = time.perf_counter()
    all_stats = gather_stats(local_stats, rank, world, dev)
    _barrier(dev)
    gather_ms = 1e3 * (time.perf_counter() - tg)

    if rank == 0:
        # batch-1 step latency (BASELINE metric, config 1): one environment, fused rollout, device time per RK4 step, for
        # the one-warp kernel (the default) and the generic 64- and 128-thread-per-environment kernels
        lat = {}
        for name, kw in (("32", dict(threads_per_env=32)), ("64", dict(threads_per_env=64)), ("128", dict(threads_per_env=128)),
                         ("32+2 helper warps (latency mode)", dict(threads_per_env=32, envs_per_block=-1))):
            b1 = lmh.Batch(1, device=local_rank, **kw)
            b1.step(20)
            lat[name] = min(b1.step_timed(200) for _ in range(3)) / 200 * 1e3
            b1.close()
        lat_us = lat["32+2 helper warps (latency mode)"]
        fp64_tf, sm_mhz_attr, sms = lmh.measure_fp64_peak(local_rank)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        total_env_steps = world * E * S * K
        value = total_env_steps / (ms_total * 1e-3)
        # end to end = host copies + rollouts + (N > 1) the final gather of the statistics, once per K rollouts' worth of work
        e2e_value = total_env_steps / ((e2e_ms + (K * gather_ms if world > 1 else 0.0)) * 1e-3)
        ach_tf = FLOP_PER_ENV_STEP * E * S / (kern * 1e-3) / 1e12
        ach_gbs = STATE_BYTES_PER_ENV * E / (kern * 1e-3) / 1e9
        info = b.launch_info()
        line = {
            "metric": "closed-loop env-steps/sec (NAO, FP64)", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": K, "warmup": max(W, 3), "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload],
                       "envs_per_gpu": E, "rk4_steps_per_rollout": S, "stage_evals_per_env_step": 4, "seed": hex(seed).upper().replace("X", "x"),
                       "push": "U(0,0.25) m/s on base v_xy at step U{0..50}" if args.workload != "walk" else "none",
                       "mpc_horizon": 160 if walk else 50, "dt": 0.01,
                       "threads_per_env": info["threads_per_env"], "smem_bytes_per_env": info["smem_bytes"],
                       "l2": "flushed (256 MiB write) before every timed rollout; working set is on-chip"},
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / K + (gather_ms if world > 1 else 0.0),
                    "includes_gather_ms": gather_ms if world > 1 else 0.0},
            "gpu_launches": launches,
            "batch1_step_latency_us": lat_us, "batch1_step_latency_us_by_threads_per_env": lat,
            "wall_ms_timed_region": wall_ms,
            "gather_ms": gather_ms,
            "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": fp64_tf, "unit": "TFLOP/s", "frac": ach_tf / fp64_tf,
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH[args.workload], "traffic_source": NCU_TRAFFIC_SOURCE[args.workload], "kernel": "rollout_kernel_multi" if info["envs_per_block"] > 1 else "rollout_kernel", "kernel_ms_per_launch": kern,
                         "flop_per_env_step": FLOP_PER_ENV_STEP,
                         "peak_source": "BUILDER-MEASURED: FP64 FMA micro-kernel run in this process (lmh_measure_fp64_peak; "
                                        "97 % of the nominal 148 SM x 64 FMA x 2 x 1.965 GHz = 37.2 TFLOP/s); "
                                        "MEASURED_PEAKS.json has no FP64 entry and the profiling guide states no FP64 fallback",
                         "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback"}},
            "clocks": clocks,
            "rollout_stats": {"alive_frac": float(all_stats[:, 6].mean()), "qp_fail": int(all_stats[:, 4].sum()),
                              "qp_iters_per_stage": float(all_stats[:, 3].sum() / (all_stats[:, 5].sum() * 4.0)),
                              "max_abs_tau": float(all_stats[:, 0].max()), "max_com_err": float(all_stats[:, 1].max())},
        }
        if world == 1 and not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import orc_py
            orc_py.build()
            lib = orc_py.oracle()
            cores = os.cpu_count() or 1
            # a bounded sample of the SAME workload: the first n_cpu envs, full rollout length (the 1.6 s horizon makes an
            # oracle stage about three times dearer, hence fewer envs for the walking workloads)
            n_cpu, s_cpu = (32 if not walk else 12) * cores, S      # about 10 s of CPU work on the box's 16 threads
            wp = lmh.walk_plan(range(n_cpu), seed=seed) if walk else None
            ps, dv = push_step[:n_cpu], push_dv[:n_cpu]
            cfg = lib.default_config(second_wbc=0, mpc_rebuild_hessian=0, **({"time_horizon": 1.6} if walk else {}))
            t0 = time.perf_counter()
            r = lib.batch_rollout_ex(cfg, n_cpu, s_cpu, push_step=ps, push_dv=dv, walk=wp, n_threads=cores)
            dtc = time.perf_counter() - t0
            so = r["stats"]
            rate = float(so["steps"].sum()) / dtc
            # the same envs on the GPU for the same number of steps: a live parity check beside the timing
            b2 = lmh.Batch(n_cpu, device=local_rank, threads_per_env=args.threads_per_env, **({"time_horizon": 1.6} if walk else {}))
            if walk:
                b2.make_walk_plans(first_env_id=0, seed=seed)
            b2.set_pushes(ps, dv)
            b2.step(s_cpu)
            xg, _, _ = b2.get_state()
            sg = b2.stats()
            b2.close()
            both = (sg["alive"] == 1) & (so["alive"] == 1)
            line["cpu_baseline"] = {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                    "sample": f"first {n_cpu} envs of the workload x {s_cpu} RK4 steps, {cores} host threads, {dtc:.1f} s",
                                    "note": "oracle port: one WBC solve per stage, cached MPC factor, -O2, no sanitizers",
                                    "max_abs_state_diff_vs_gpu": float(np.abs(xg - r["x"])[both].max()) if both.any() else None,
                                    "integer_stats_equal": bool(np.array_equal(sg["alive"], so["alive"]) and np.array_equal(sg["steps"], so["steps"])
                                                                and np.array_equal(sg["qp_fail"], so["qp_fail"])),
                                    "alive_in_sample": int(so["alive"].sum())}
        _emit(line)
    b.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _emit(line):
    """the ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) was sent to stderr"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                   # libraries that print to stdout (NCCL version banner) must not pollute the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=0, help="environments per GPU (default: the workload's BASELINE size)")
    ap.add_argument("--rollout", type=int, default=0, help="RK4 steps per rollout = per bench step (default: the workload's)")
    ap.add_argument("--threads-per-env", type=int, default=0)
    ap.add_argument("--workload", default="push", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.envs = args.envs or DEFAULTS[args.workload][0]
    args.rollout = args.rollout or DEFAULTS[args.workload][1]
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_qp_microbench(args) if args.workload == "qp" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
