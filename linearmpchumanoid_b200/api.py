"""ctypes binding of liblmh_b200.so (include/lmh_capi.h).  No compute happens in Python and there is
no CPU fallback: if the shared library or a CUDA device is missing, construction raises."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "liblmh_b200.so")
_f64, _i32 = C.c_double, C.c_int32

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]


class LmhError(RuntimeError):
    pass


class LmhConfig(C.Structure):
    _fields_ = [("dt", _f64), ("sim_time", _f64), ("time_horizon", _f64), ("com", _f64 * 3), ("rf", _f64 * 6),
                ("lf", _f64 * 6), ("mu", _f64), ("mpc_alpha", _f64), ("mpc_beta", _f64), ("qp_max_iter", _i32),
                ("threads_per_env", _i32), ("device", _i32), ("log_steps", _i32), ("envs_per_block", _i32), ("placement", _i32), ("mpc_zcom", _f64)]


class LmhWalkSpec(C.Structure):
    _fields_ = [("n_footsteps", _i32), ("horizon_steps", _i32), ("tail_steps", _i32), ("pad", _i32), ("time_per_step", _f64),
                ("dt", _f64), ("step_height", _f64), ("lateral", _f64), ("min_len", _f64), ("max_len", _f64), ("fixed_len", _f64)]


STAGE_DTYPE = np.dtype([("M", "f8", (30, 30)), ("C", "f8", 30), ("Cg", "f8", 30), ("AG", "f8", (6, 30)),
                        ("AGpqp", "f8", 6), ("Jpqp", "f8", 12), ("JFeet", "f8", (12, 30)), ("CoM", "f8", 3),
                        ("comVel", "f8", 3), ("comAngMom", "f8", 3), ("xRef", "f8", 3), ("yRef", "f8", 3),
                        ("qppRef", "f8", 30), ("hGpRef", "f8", 6), ("footAccRef", "f8", 12), ("Ha", "f8", (30, 30)),
                        ("ga", "f8", 30), ("a", "f8", 30), ("f", "f8", 12), ("c", "f8", 32), ("lam", "f8", 6),
                        ("qpp", "f8", 30), ("tau", "f8", 24), ("T7", "f8", 12), ("T14", "f8", 12), ("xdot", "f8", 60),
                        ("k_index", "i4"), ("qp_status", "i4"), ("qp_iters", "i4"), ("phase", "i4")])
STEPLOG_DTYPE = np.dtype([("com_k4", "f8", 3), ("rfoot", "f8", 3), ("lfoot", "f8", 3), ("tau_k1", "f8", 24),
                          ("k_index", "i4", 4), ("qp_iters", "i4"), ("qp_fail", "i4"), ("phase", "i4", 4),
                          ("contact", "i4", 4)])
STATS_DTYPE = np.dtype([("max_abs_tau", "f8"), ("max_com_err", "f8"), ("max_abs_qpp", "f8"), ("qp_iters", "i4"),
                        ("qp_fail", "i4"), ("steps", "i4"), ("alive", "i4"), ("qp_fallback", "i4"), ("pad", "i4")])

# health rule of a closed-loop step (include/lmh_capi.h: LMH_TAU_LIMIT, LMH_COM_ERR_LIMIT, LMH_STATE_LIMIT, LMH_TORSO_MIN)
TAU_LIMIT, COM_ERR_LIMIT, STATE_LIMIT, TORSO_MIN = 100.0, 0.5, 1e6, 0.15

# every symbol include/lmh_capi.h declares (checked by the CPU test-suite against the built library)
SYMBOLS = ["lmh_last_error", "lmh_default_config", "lmh_create", "lmh_destroy", "lmh_n_env", "lmh_get_nominal",
           "lmh_get_mpc_gain", "lmh_get_link_tables", "lmh_set_plans", "lmh_set_walk_plans", "lmh_get_plan", "lmh_set_state",
           "lmh_get_state", "lmh_reset", "lmh_set_pushes", "lmh_set_step_counter", "lmh_stage", "lmh_step",
           "lmh_step_timed", "lmh_sync", "lmh_get_stats", "lmh_get_log", "lmh_get_com", "lmh_state_dev",
           "lmh_stats_dev", "lmh_set_state_dev", "lmh_reset_stats", "lmh_timer_start", "lmh_timer_stop",
           "lmh_qp_bench", "lmh_ik_batch", "lmh_export_mujoco", "lmh_launch_info", "lmh_measure_fp64_peak",
           "lmh_ik_solve", "lmh_kinematics", "lmh_mpc_solve", "lmh_get_mpc_tables", "lmh_foot_coeff", "lmh_push_plan",
           "lmh_default_walk_spec", "lmh_walk_plan_dims", "lmh_walk_plan", "lmh_make_walk_plans", "lmh_get_link_inertia",
           "lmh_get_nao_parameters", "lmh_local_transforms", "lmh_poly_fit"]


def build_library(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... in-tree (cross-compiles without a GPU)."""
    src = os.path.join(_HERE, "csrc", "lmh_capi.cu")
    deps = [src] + [os.path.join(_HERE, "csrc", f) for f in ("lmh_model.cuh", "lmh_warp.cuh", "lmh_ik.cuh", "lmh_setup.hpp", "lmh_nao_table.inc")]
    deps.append(os.path.join(_ROOT, "include", "lmh_capi.h"))
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + ["-I", os.path.join(_ROOT, "include"), "-o", LIB_PATH, src]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise LmhError("nvcc failed:\n" + r.stdout + r.stderr)
    return r.stderr if verbose else LIB_PATH


_lib = None


def load_library():
    """Loads liblmh_b200.so; raises if it was never built (there is nothing to fall back to)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LmhError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.lmh_last_error.restype = C.c_char_p
    L.lmh_default_config.argtypes = [C.POINTER(LmhConfig)]
    L.lmh_create.argtypes = [C.POINTER(LmhConfig), _i32, C.POINTER(vp)]
    L.lmh_destroy.argtypes = [vp]
    L.lmh_n_env.argtypes = [vp]
    L.lmh_get_nominal.argtypes = [vp, vp, vp, vp, vp]
    L.lmh_get_mpc_gain.argtypes = [vp, vp, _i32, vp, vp]
    L.lmh_get_link_tables.argtypes = [vp, vp, vp, vp]
    L.lmh_set_plans.argtypes = [vp, _i32, _i32, vp, vp, vp, vp]
    L.lmh_set_walk_plans.argtypes = [vp, _i32, _i32, vp, vp, _i32, _f64, vp, vp, vp]
    L.lmh_get_plan.argtypes = [vp, _i32, vp, vp, _i32, vp, vp]
    L.lmh_set_state.argtypes = [vp, vp, vp, vp]
    L.lmh_get_state.argtypes = [vp, vp, vp, vp]
    L.lmh_reset.argtypes = [vp]
    L.lmh_set_pushes.argtypes = [vp, vp, vp]
    L.lmh_set_step_counter.argtypes = [vp, _i32]
    L.lmh_stage.argtypes = [vp, vp]
    L.lmh_step.argtypes = [vp, _i32]
    L.lmh_step_timed.argtypes = [vp, _i32, vp]
    L.lmh_sync.argtypes = [vp]
    L.lmh_get_stats.argtypes = [vp, vp]
    L.lmh_get_log.argtypes = [vp, _i32, _i32, vp]
    L.lmh_get_com.argtypes = [vp, vp]
    L.lmh_state_dev.argtypes = [vp, vp, vp, vp, vp]
    L.lmh_stats_dev.argtypes = [vp, vp]
    L.lmh_set_state_dev.argtypes = [vp, vp, vp, vp]
    L.lmh_reset_stats.argtypes = [vp]
    L.lmh_timer_start.argtypes = [vp]
    L.lmh_timer_stop.argtypes = [vp, vp]
    L.lmh_qp_bench.argtypes = [vp, _i32, _i32, vp, vp]
    L.lmh_ik_batch.argtypes = [vp, _i32, vp, vp, vp, vp, vp, vp, vp, _i32]
    L.lmh_export_mujoco.argtypes = [_i32, vp, vp]
    L.lmh_launch_info.argtypes = [vp, vp, vp, vp, vp, vp]
    L.lmh_measure_fp64_peak.argtypes = [_i32, vp, vp, vp]
    L.lmh_ik_solve.argtypes = [vp, _i32, vp, vp, vp, vp, vp, vp]
    L.lmh_kinematics.argtypes = [vp, _i32, vp, vp, vp, vp]
    L.lmh_mpc_solve.argtypes = [vp, _i32, vp, vp, vp, vp, vp, vp, _i32, _i32, vp]
    L.lmh_get_mpc_tables.argtypes = [vp, vp, vp, _i32, vp]
    L.lmh_foot_coeff.argtypes = [vp, vp, _f64, _f64, vp, vp, vp]
    L.lmh_push_plan.argtypes = [vp, _i32, C.c_uint64, _f64, _i32, vp, vp]
    L.lmh_default_walk_spec.argtypes = [C.POINTER(LmhWalkSpec)]
    L.lmh_default_walk_spec.restype = None
    L.lmh_walk_plan_dims.argtypes = [C.POINTER(LmhWalkSpec), vp, vp]
    L.lmh_walk_plan.argtypes = [C.POINTER(LmhWalkSpec), vp, _i32, C.c_uint64, vp, vp, vp, vp, vp]
    L.lmh_make_walk_plans.argtypes = [vp, C.POINTER(LmhWalkSpec), C.c_int64, C.c_uint64]
    L.lmh_get_link_inertia.argtypes = [vp, vp]
    L.lmh_get_nao_parameters.argtypes = [vp, vp, vp]
    L.lmh_local_transforms.argtypes = [vp, vp]
    L.lmh_poly_fit.argtypes = [_i32, vp, _i32, vp, _i32, vp, vp]
    _lib = L
    return L


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _chk(L, rc):
    if rc != 0:
        raise LmhError(f"lmh error {rc}: {L.lmh_last_error().decode()}")


def measure_fp64_peak(device=0):
    L = load_library()
    tf, mhz, sms = np.zeros(1), np.zeros(1), np.zeros(1, np.int32)
    _chk(L, L.lmh_measure_fp64_peak(device, _p(tf), _p(mhz), _p(sms)))
    return float(tf[0]), float(mhz[0]), int(sms[0])


def export_mujoco(q):
    """relabelMujocoMatrix + initial free-joint pose (apps/mujoco/main.cpp:182-200, MujocoSim.cpp:77-85):
    q [n][30] -> qpos [n][31] = [x, y, z + 0.035, quat wxyz = identity, 24 joints in MuJoCo's order]"""
    L = load_library()
    q = np.ascontiguousarray(q, np.float64).reshape(-1, 30)
    out = np.zeros((q.shape[0], 31))
    _chk(L, L.lmh_export_mujoco(q.shape[0], _p(q), _p(out)))
    return out


def walk_spec(**kw):
    L = load_library()
    ws = LmhWalkSpec()
    L.lmh_default_walk_spec(C.byref(ws))
    for k, v in kw.items():
        if not hasattr(ws, k):
            raise LmhError(f"unknown walk-spec field {k}")
        setattr(ws, k, v)
    return ws


class Batch:
    """n_env independent NAO environments stepped in lockstep on one GPU.

    Mirrors what apps/offline/main.cpp builds (Robot + IK + Mpc3dLip + ZMP + foot polynomials +
    Controller) and what its stand() loop does (rk4Step + Clock::step), for a whole batch."""

    def __init__(self, n_env, device=0, threads_per_env=0, log_steps=0, envs_per_block=0, placement=0, **scenario):
        self.L = L = load_library()
        cfg = LmhConfig()
        L.lmh_default_config(C.byref(cfg))
        cfg.device, cfg.threads_per_env, cfg.log_steps = device, threads_per_env, log_steps
        cfg.envs_per_block, cfg.placement = envs_per_block, placement
        for k, v in scenario.items():
            if not hasattr(cfg, k):
                raise LmhError(f"unknown scenario field {k}")
            if isinstance(v, (list, tuple, np.ndarray)):
                arr = getattr(cfg, k)
                for i, x in enumerate(v):
                    arr[i] = x
            else:
                setattr(cfg, k, v)
        self.cfg = cfg
        self.n_env = int(n_env)
        self.h = C.c_void_p()
        _chk(L, L.lmh_create(C.byref(cfg), self.n_env, C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None):
            self.L.lmh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setup results
    def nominal(self):
        q0, z, m, it = np.zeros(30), np.zeros(1), np.zeros(1), np.zeros(1, np.int32)
        _chk(self.L, self.L.lmh_get_nominal(self.h, _p(q0), _p(z), _p(m), _p(it)))
        return q0, float(z[0]), float(m[0]), int(it[0])

    def mpc_gain(self):
        r, n1, rpx = np.zeros(512), np.zeros(1, np.int32), np.zeros(2)
        _chk(self.L, self.L.lmh_get_mpc_gain(self.h, _p(r), 512, _p(n1), _p(rpx)))
        return r[:int(n1[0])].copy(), rpx

    def link_tables(self):
        m, c, i = np.zeros(28), np.zeros((28, 3)), np.zeros((28, 6, 6))
        _chk(self.L, self.L.lmh_get_link_tables(self.h, _p(m), _p(c), _p(i)))
        return m, c, i

    def plan(self, k=0):
        n = np.zeros(1, np.int32)
        _chk(self.L, self.L.lmh_get_plan(self.h, k, None, None, 0, _p(n), None))
        zx, zy, fc = np.zeros(int(n[0])), np.zeros(int(n[0])), np.zeros(40)
        _chk(self.L, self.L.lmh_get_plan(self.h, k, _p(zx), _p(zy), int(n[0]), _p(n), _p(fc)))
        return zx, zy, fc.reshape(2, 20)

    def set_plans(self, zmp_x, zmp_y, foot_coef, plan_ids=None):
        zx = np.ascontiguousarray(np.atleast_2d(zmp_x), np.float64)
        zy = np.ascontiguousarray(np.atleast_2d(zmp_y), np.float64)
        fc = np.ascontiguousarray(foot_coef, np.float64).reshape(zx.shape[0], 40)
        ids = None if plan_ids is None else np.ascontiguousarray(plan_ids, np.int32)
        if zy.shape != zx.shape or (ids is not None and ids.size != self.n_env):
            raise LmhError("set_plans: zmp_x / zmp_y shapes differ or plan_ids is not one id per environment")
        _chk(self.L, self.L.lmh_set_plans(self.h, zx.shape[0], zx.shape[1], _p(zx), _p(zy), _p(fc), _p(ids)))

    def set_walk_plans(self, zmp_x, zmp_y, phase_T, foot_coef, contact, plan_ids=None):
        """EXTENSION (walking): zmp_x/zmp_y [n_plans][len]; foot_coef [n_plans][n_phases][2][20]; contact [n_plans][n_phases]"""
        zx = np.ascontiguousarray(np.atleast_2d(zmp_x), np.float64)
        zy = np.ascontiguousarray(np.atleast_2d(zmp_y), np.float64)
        ct = np.ascontiguousarray(contact, np.int32).reshape(zx.shape[0], -1)
        fc = np.ascontiguousarray(foot_coef, np.float64).reshape(zx.shape[0], ct.shape[1], 40)
        ids = None if plan_ids is None else np.ascontiguousarray(plan_ids, np.int32)
        if zy.shape != zx.shape or (ids is not None and ids.size != self.n_env):
            raise LmhError("set_walk_plans: zmp_x / zmp_y shapes differ or plan_ids is not one id per environment")
        _chk(self.L, self.L.lmh_set_walk_plans(self.h, zx.shape[0], zx.shape[1], _p(zx), _p(zy), ct.shape[1], float(phase_T),
                                               _p(fc), _p(ct), _p(ids)))

    # ---- state
    def reset(self):
        _chk(self.L, self.L.lmh_reset(self.h))

    def set_state(self, x=None, v_held=None, t=None):
        a = [None if v is None else np.ascontiguousarray(v, np.float64) for v in (x, v_held, t)]
        for arr, n in zip(a, (60, 30, 1)):
            if arr is not None and arr.size != self.n_env * n:
                raise LmhError("set_state: wrong array size")
        _chk(self.L, self.L.lmh_set_state(self.h, _p(a[0]), _p(a[1]), _p(a[2])))
        _chk(self.L, self.L.lmh_sync(self.h))

    def get_state(self):
        x, vh, t = np.zeros((self.n_env, 60)), np.zeros((self.n_env, 30)), np.zeros(self.n_env)
        _chk(self.L, self.L.lmh_get_state(self.h, _p(x), _p(vh), _p(t)))
        return x, vh, t

    def set_pushes(self, push_step, push_dv6):
        ps = None if push_step is None else np.ascontiguousarray(push_step, np.int32)
        dv = None if push_dv6 is None else np.ascontiguousarray(push_dv6, np.float64)
        if (ps is not None and ps.size != self.n_env) or (dv is not None and dv.size != 6 * self.n_env):
            raise LmhError("set_pushes: need n_env push steps and n_env x 6 impulses")
        _chk(self.L, self.L.lmh_set_pushes(self.h, _p(ps), _p(dv)))

    def set_step_counter(self, step):
        _chk(self.L, self.L.lmh_set_step_counter(self.h, int(step)))

    # ---- compute
    def stage(self, dump=True):
        d = np.zeros(self.n_env, STAGE_DTYPE) if dump else None
        _chk(self.L, self.L.lmh_stage(self.h, _p(d)))
        return d

    def step(self, n_steps, sync=True):
        _chk(self.L, self.L.lmh_step(self.h, int(n_steps)))
        if sync:
            _chk(self.L, self.L.lmh_sync(self.h))

    def step_timed(self, n_steps):
        ms = np.zeros(1)
        _chk(self.L, self.L.lmh_step_timed(self.h, int(n_steps), _p(ms)))
        return float(ms[0])

    def sync(self):
        _chk(self.L, self.L.lmh_sync(self.h))

    def stats(self):
        s = np.zeros(self.n_env, STATS_DTYPE)
        _chk(self.L, self.L.lmh_get_stats(self.h, _p(s)))
        return s

    def log(self, first=0, n=None):
        n = self.cfg.log_steps - first if n is None else n
        out = np.zeros((n, self.n_env), STEPLOG_DTYPE)
        _chk(self.L, self.L.lmh_get_log(self.h, first, n, _p(out)))
        return out

    def com(self):
        c = np.zeros((self.n_env, 3))
        _chk(self.L, self.L.lmh_get_com(self.h, _p(c)))
        return c

    def qp_bench(self, mode, repeats):
        """config 5: (ms, a[n_env][30], f[n_env][12]) of `repeats` WBC QP solves (mode 0) / MPC evaluations (mode 1)"""
        out, ms = np.zeros((self.n_env, 42)), np.zeros(1)
        _chk(self.L, self.L.lmh_qp_bench(self.h, int(mode), int(repeats), _p(out), _p(ms)))
        return float(ms[0]), out[:, :30], out[:, 30:]

    def ik(self, com, rf, lf, apply=False):
        """Kinematics::compute for n targets on the device (invKinematics.cpp:11-52): com [n][3], rf / lf [n][6] (sole
        position + Euler angles).  Returns {"q": [n][30], "iters", "residual", "status"}; apply=True (n == n_env)
        restarts every environment from its own posture."""
        com = np.ascontiguousarray(com, np.float64).reshape(-1, 3)
        n = com.shape[0]
        rf = np.ascontiguousarray(np.broadcast_to(np.asarray(rf, np.float64).reshape(-1, 6), (n, 6)))
        lf = np.ascontiguousarray(np.broadcast_to(np.asarray(lf, np.float64).reshape(-1, 6), (n, 6)))
        q, it, res, st = np.zeros((n, 30)), np.zeros(n, np.int32), np.zeros(n), np.zeros(n, np.int32)
        _chk(self.L, self.L.lmh_ik_batch(self.h, n, _p(com), _p(rf), _p(lf), _p(q), _p(it), _p(res), _p(st), 1 if apply else 0))
        return {"q": q, "iters": it, "residual": res, "status": st}

    def ik_solve(self, desop, q_start=None):
        """general Kinematics::compute: desop [n][30] operational targets, q_start [n][30] or None (nominal posture)"""
        des = np.ascontiguousarray(desop, np.float64).reshape(-1, 30)
        n = des.shape[0]
        qs = None if q_start is None else np.ascontiguousarray(np.broadcast_to(np.asarray(q_start, np.float64).reshape(-1, 30), (n, 30)))
        q, it, res, st = np.zeros((n, 30)), np.zeros(n, np.int32), np.zeros(n), np.zeros(n, np.int32)
        _chk(self.L, self.L.lmh_ik_solve(self.h, n, _p(qs), _p(des), _p(q), _p(it), _p(res), _p(st)))
        return {"q": q, "iters": it, "residual": res, "status": st}

    def kinematics(self, q):
        """Robot::updateState for n postures: (T [n][28][4][4], X [n][28][6][6], com [n][3])"""
        q = np.ascontiguousarray(q, np.float64).reshape(-1, 30)
        n = q.shape[0]
        T, X, com = np.zeros((n, 28, 4, 4)), np.zeros((n, 28, 6, 6)), np.zeros((n, 3))
        _chk(self.L, self.L.lmh_kinematics(self.h, n, _p(q), _p(T), _p(X), _p(com)))
        return T, X, com

    def mpc_tables(self):
        n1 = np.zeros(1, np.int32)
        _chk(self.L, self.L.lmh_get_mpc_tables(self.h, None, None, 0, _p(n1)))
        n = int(n1[0])
        Px, Pu = np.zeros((n, 2)), np.zeros((n, n))
        _chk(self.L, self.L.lmh_get_mpc_tables(self.h, _p(Px), _p(Pu), n, _p(n1)))
        return Px, Pu

    def mpc_solve(self, com_xy, vel_xy, zmp_x, zmp_y, use_dmma=False, repeats=1):
        """Mpc3dLip::compute as a dense solve per problem on the device: returns (xref [n][3], yref [n][3], ms)"""
        c = np.ascontiguousarray(com_xy, np.float64).reshape(-1, 2)
        n = c.shape[0]
        v = np.ascontiguousarray(vel_xy, np.float64).reshape(n, 2)
        zx = np.ascontiguousarray(zmp_x, np.float64).reshape(n, -1)
        zy = np.ascontiguousarray(zmp_y, np.float64).reshape(n, -1)
        n1 = np.zeros(1, np.int32)
        _chk(self.L, self.L.lmh_get_mpc_tables(self.h, None, None, 0, _p(n1)))
        if zx.shape[1] != int(n1[0]) or zy.shape != zx.shape:
            raise LmhError("mpc_solve: the ZMP windows must hold horizon + 1 samples")
        xr, yr, ms = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(1)
        _chk(self.L, self.L.lmh_mpc_solve(self.h, n, _p(c), _p(v), _p(zx), _p(zy), _p(xr), _p(yr), 1 if use_dmma else 0, int(repeats), _p(ms)))
        return xr, yr, float(ms[0])

    def make_walk_plans(self, first_env_id=0, seed=0x5EED0003, **spec):
        """generate and install one walking plan per environment inside the library (no trip through Python)"""
        ws = walk_spec(**spec)
        _chk(self.L, self.L.lmh_make_walk_plans(self.h, C.byref(ws), int(first_env_id), int(seed)))

    def launch_info(self):
        v = [np.zeros(1, np.int32) for _ in range(5)]
        _chk(self.L, self.L.lmh_launch_info(self.h, *[_p(a) for a in v]))
        return dict(zip(("threads_per_env", "envs_per_block", "smem_bytes", "blocks", "launches"),
                        (int(a[0]) for a in v)))

    def device_pointers(self):
        x, vh, t, s = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        _chk(self.L, self.L.lmh_state_dev(self.h, C.byref(x), C.byref(vh), C.byref(t), C.byref(s)))
        return x.value, vh.value, t.value, s.value

    def set_state_dev(self, x_ptr=None, vh_ptr=None, t_ptr=None):
        """device-to-device restore from caller-owned device buffers (e.g. torch tensors' data_ptr())"""
        _chk(self.L, self.L.lmh_set_state_dev(self.h, C.c_void_p(x_ptr) if x_ptr else None,
                                              C.c_void_p(vh_ptr) if vh_ptr else None, C.c_void_p(t_ptr) if t_ptr else None))

    def reset_stats(self):
        _chk(self.L, self.L.lmh_reset_stats(self.h))

    def timer_start(self):
        _chk(self.L, self.L.lmh_timer_start(self.h))

    def timer_stop(self):
        ms = np.zeros(1)
        _chk(self.L, self.L.lmh_timer_stop(self.h, _p(ms)))
        return float(ms[0])

    # raw-pointer variants for pinned host buffers (used by the end-to-end bench leg)
    def set_state_ptr(self, x_ptr, vh_ptr=None, t_ptr=None):
        _chk(self.L, self.L.lmh_set_state(self.h, C.c_void_p(x_ptr), C.c_void_p(vh_ptr) if vh_ptr else None,
                                          C.c_void_p(t_ptr) if t_ptr else None))

    def get_state_ptr(self, x_ptr):
        _chk(self.L, self.L.lmh_get_state(self.h, C.c_void_p(x_ptr), None, None))

    def get_stats_ptr(self, s_ptr):
        _chk(self.L, self.L.lmh_get_stats(self.h, C.c_void_p(s_ptr)))
