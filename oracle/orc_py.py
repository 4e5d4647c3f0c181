"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes view of oracle/liblmh_oracle.so
(the CPU restatement) and oracle/_ref/liblmh_ref.so (the reference's own sources
built against shim headers).  Import only from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
f64 = C.c_double
i32 = C.c_int32


class OrcConfig(C.Structure):
    _fields_ = [("dt", f64), ("sim_time", f64), ("time_horizon", f64), ("com", f64 * 3), ("rf", f64 * 6),
                ("lf", f64 * 6), ("second_wbc", i32), ("mpc_rebuild_hessian", i32), ("qp_long_double", i32),
                ("reserved", i32)]


_DUMP = [("M", 900), ("C", 30), ("Cg", 30), ("AG", 180), ("AGpqp", 6), ("Jpqp", 12), ("JFeet", 360),
         ("CoM", 3), ("comVel", 3), ("comAngMom", 3), ("xRef", 3), ("yRef", 3), ("qppRef", 30), ("hGpRef", 6),
         ("footAccRef", 12), ("H", 5476), ("g", 74), ("x", 74), ("y", 50), ("qpp", 30), ("tau", 24), ("f", 12),
         ("T7", 16), ("T14", 16), ("kkt", 4), ("xdot", 60)]
_SHAPES = {"M": (30, 30), "AG": (6, 30), "JFeet": (12, 30), "H": (74, 74), "T7": (4, 4), "T14": (4, 4)}


class OrcStageDump(C.Structure):
    _fields_ = [(n, f64 * k) for n, k in _DUMP] + [("k_index", i32), ("qp_status", i32), ("qp_iters", i32),
                                                   ("pad", i32)]

    def to_dict(self):
        d = {}
        for n, k in _DUMP:
            a = np.ctypeslib.as_array(getattr(self, n)).copy()
            d[n] = a.reshape(_SHAPES[n]) if n in _SHAPES else a
        d["k_index"], d["qp_status"], d["qp_iters"] = self.k_index, self.qp_status, self.qp_iters
        return d


class OrcStepLog(C.Structure):
    _fields_ = [("x", f64 * 60), ("com_k4", f64 * 3), ("rfoot", f64 * 3), ("lfoot", f64 * 3), ("tau_k1", f64 * 24),
                ("k_index", i32 * 4), ("qp_iters", i32), ("qp_fail", i32), ("phase", i32 * 4), ("contact", i32 * 4)]


class OrcWalkPlans(C.Structure):
    _fields_ = [("zmp_len", i32), ("n_phases", i32), ("phase_T", f64), ("zmp_x", C.c_void_p), ("zmp_y", C.c_void_p),
                ("foot_coef", C.c_void_p), ("contact", C.c_void_p)]


STEPLOG_DTYPE = np.dtype([("x", "f8", 60), ("com_k4", "f8", 3), ("rfoot", "f8", 3), ("lfoot", "f8", 3),
                          ("tau_k1", "f8", 24), ("k_index", "i4", 4), ("qp_iters", "i4"), ("qp_fail", "i4"),
                          ("phase", "i4", 4), ("contact", "i4", 4)])


def build(ref=False):
    """Compile the checker (g++ only).  ref=True also builds oracle/_ref when /root/reference exists."""
    subprocess.run(["make", "-C", HERE, "liblmh_oracle.so"], check=True, capture_output=True)
    if ref and os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Lib:
    def __init__(self, path):
        self.path = path
        L = self.L = C.CDLL(path)
        L.orc_sim_create.restype = C.c_void_p
        L.orc_sim_create.argtypes = [C.POINTER(OrcConfig)]
        L.orc_sim_destroy.argtypes = [C.c_void_p]
        L.orc_sim_ik_iterations.argtypes = [C.c_void_p]
        L.orc_sim_mass.restype = f64
        L.orc_sim_mass.argtypes = [C.c_void_p]
        L.orc_sim_zcom.restype = f64
        L.orc_sim_zcom.argtypes = [C.c_void_p]
        L.orc_sim_get_state.argtypes = [C.c_void_p] * 4
        L.orc_sim_set_state.argtypes = [C.c_void_p, C.c_void_p, f64, C.c_void_p]
        L.orc_sim_stage.argtypes = [C.c_void_p, C.c_void_p, f64, C.POINTER(OrcStageDump)]
        L.orc_sim_step.argtypes = [C.c_void_p, i32, C.c_void_p]
        L.orc_sim_run_stand.argtypes = [C.c_void_p, C.c_void_p, i32]
        L.orc_sim_mpc_tables.argtypes = [C.c_void_p] * 4
        L.orc_sim_foot_coeffs.argtypes = [C.c_void_p] * 7
        L.orc_sim_zmp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, i32]
        L.orc_sim_link_table.argtypes = [C.c_void_p] * 4
        L.orc_batch_rollout.argtypes = [C.POINTER(OrcConfig), i32, C.c_void_p, C.c_void_p, C.c_void_p, i32, i32,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
        if hasattr(L, "orc_sim_set_walk_plan"):
            L.orc_sim_set_walk_plan.argtypes = [C.c_void_p, i32, C.c_void_p, C.c_void_p, i32, f64, C.c_void_p, C.c_void_p]
            L.orc_sim_last_phase.argtypes = [C.c_void_p]
        if hasattr(L, "orc_sim_frames"):
            L.orc_sim_frames.argtypes = [C.c_void_p] * 3
        if hasattr(L, "orc_batch_rollout_ex"):
            L.orc_batch_rollout_ex.argtypes = [C.POINTER(OrcConfig), i32, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, i32, i32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            L.orc_sim_alive.argtypes = [C.c_void_p]
            L.orc_sim_steps.argtypes = [C.c_void_p]
        if hasattr(L, "orc_qp_solve"):
            L.orc_qp_solve.argtypes = [i32, C.c_void_p, C.c_void_p, i32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, i32, C.c_void_p, C.c_void_p]

    def default_config(self, **kw):
        cfg = OrcConfig()
        self.L.orc_default_config(C.byref(cfg))
        for k, v in kw.items():
            if isinstance(v, (list, tuple, np.ndarray)):
                arr = getattr(cfg, k)
                for i, x in enumerate(v):
                    arr[i] = x
            else:
                setattr(cfg, k, v)
        return cfg

    def sim(self, cfg=None, **kw):
        return Sim(self, cfg or self.default_config(**kw))

    def batch_rollout(self, cfg, n_env, n_steps, x0=None, push_step=None, push_dv=None, n_threads=1):
        x0 = None if x0 is None else np.ascontiguousarray(x0, np.float64)
        ps = None if push_step is None else np.ascontiguousarray(push_step, np.int32)
        pd = None if push_dv is None else np.ascontiguousarray(push_dv, np.float64)
        x = np.zeros((n_env, 60))
        com = np.zeros((n_env, 3))
        st = np.zeros((n_env, 4))
        fails = self.L.orc_batch_rollout(C.byref(cfg), n_env, _p(x0), _p(ps), _p(pd), n_steps, n_threads, _p(x),
                                         _p(com), _p(st))
        return x, com, st, fails

    def batch_rollout_ex(self, cfg, n_env, n_steps, x0=None, push_step=None, push_dv=None, walk=None, n_threads=1,
                         log=False):
        """rollout with optional per-env walking plans (dict of scenarios.walk_plan), the health rule of orc_capi.h and
        the full statistics: returns dict(x, com, stats (structured), log or None, qp_failures)"""
        x0 = None if x0 is None else np.ascontiguousarray(x0, np.float64)
        ps = None if push_step is None else np.ascontiguousarray(push_step, np.int32)
        pd = None if push_dv is None else np.ascontiguousarray(push_dv, np.float64)
        plans, keep = None, []
        if walk is not None:
            zx = np.ascontiguousarray(walk["zmp_x"], np.float64); zy = np.ascontiguousarray(walk["zmp_y"], np.float64)
            ct = np.ascontiguousarray(walk["contact"], np.int32).reshape(n_env, -1)
            fc = np.ascontiguousarray(walk["foot_coef"], np.float64).reshape(n_env, ct.shape[1], 40)
            keep = [zx, zy, ct, fc]
            plans = OrcWalkPlans(zx.shape[1], ct.shape[1], float(walk["phase_T"]), zx.ctypes.data, zy.ctypes.data,
                                 fc.ctypes.data, ct.ctypes.data)
        x, com, st = np.zeros((n_env, 60)), np.zeros((n_env, 3)), np.zeros((n_env, 8))
        lg = np.zeros((n_env, n_steps), STEPLOG_DTYPE) if log else None
        fails = self.L.orc_batch_rollout_ex(C.byref(cfg), n_env, _p(x0), _p(ps), _p(pd),
                                            C.byref(plans) if plans is not None else None, n_steps, n_threads, _p(x),
                                            _p(com), _p(st), _p(lg))
        del keep
        stats = dict(max_abs_tau=st[:, 0], max_com_err=st[:, 1], max_abs_qpp=st[:, 2], qp_iters=st[:, 3].astype(np.int64),
                     qp_fail=st[:, 4].astype(np.int64), steps=st[:, 5].astype(np.int64), alive=st[:, 6].astype(np.int64), t=st[:, 7])
        return dict(x=x, com=com, stats=stats, log=lg, qp_failures=fails)

    def qp_solve(self, H, g, A, lbA, ubA, long_double=False):
        H = np.ascontiguousarray(H, np.float64)
        g = np.ascontiguousarray(g, np.float64)
        n = g.size
        m = 0 if A is None else A.shape[0]
        A = np.zeros((1, n)) if A is None else np.ascontiguousarray(A, np.float64)
        lbA = np.zeros(1) if lbA is None else np.ascontiguousarray(lbA, np.float64)
        ubA = np.zeros(1) if ubA is None else np.ascontiguousarray(ubA, np.float64)
        x = np.zeros(n)
        y = np.zeros(max(m, 1))
        it = np.zeros(1, np.int32)
        kkt = np.zeros(4)
        st = self.L.orc_qp_solve(n, _p(H), _p(g), m, _p(A), _p(lbA), _p(ubA), _p(x), _p(y), int(long_double), _p(it),
                                 _p(kkt))
        return x, y[:m], st, int(it[0]), kkt


class Sim:
    def __init__(self, lib, cfg):
        self.lib, self.cfg = lib, cfg
        self.h = lib.L.orc_sim_create(C.byref(cfg))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.L.orc_sim_destroy(self.h)
            self.h = None

    @property
    def mass(self):
        return self.lib.L.orc_sim_mass(self.h)

    @property
    def zcom(self):
        return self.lib.L.orc_sim_zcom(self.h)

    @property
    def ik_iterations(self):
        return self.lib.L.orc_sim_ik_iterations(self.h)

    def get_state(self):
        x, t, v = np.zeros(60), np.zeros(1), np.zeros(30)
        self.lib.L.orc_sim_get_state(self.h, _p(x), _p(t), _p(v))
        return x, float(t[0]), v

    def set_state(self, x, t, v_held=None):
        x = np.ascontiguousarray(x, np.float64)
        v = None if v_held is None else np.ascontiguousarray(v_held, np.float64)
        self.lib.L.orc_sim_set_state(self.h, _p(x), float(t), _p(v))

    def stage(self, x, t):
        x = np.ascontiguousarray(x, np.float64)
        d = OrcStageDump()
        self.lib.L.orc_sim_stage(self.h, _p(x), float(t), C.byref(d))
        return d.to_dict()

    def step(self, n, log=True):
        logs = (OrcStepLog * n)() if log else None
        self.lib.L.orc_sim_step(self.h, n, logs)
        if not log:
            return None
        return np.frombuffer(logs, dtype=STEPLOG_DTYPE).copy()

    def run_stand(self, max_steps=100000):
        out = np.zeros(max_steps)
        n = self.lib.L.orc_sim_run_stand(self.h, _p(out), max_steps)
        return out[:n]

    def set_walk_plan(self, zx, zy, phase_T, foot_coef, contact):
        """EXTENSION (walking): foot_coef [n_phases][2][20], contact [n_phases] (bit 0 right, bit 1 left)"""
        zx = np.ascontiguousarray(zx, np.float64); zy = np.ascontiguousarray(zy, np.float64)
        fc = np.ascontiguousarray(foot_coef, np.float64); ct = np.ascontiguousarray(contact, np.int32)
        self.lib.L.orc_sim_set_walk_plan(self.h, len(zx), _p(zx), _p(zy), len(ct), float(phase_T), _p(fc), _p(ct))

    def mpc_tables(self):
        hz = np.zeros(1, np.int32)
        self.lib.L.orc_sim_mpc_tables(self.h, None, None, _p(hz))
        n = int(hz[0]) + 1
        Px, Pu = np.zeros((n, 2)), np.zeros((n, n))
        self.lib.L.orc_sim_mpc_tables(self.h, _p(Px), _p(Pu), _p(hz))
        return Px, Pu, int(hz[0])

    def foot_coeffs(self):
        a = [np.zeros(6), np.zeros(6), np.zeros(8), np.zeros(6), np.zeros(6), np.zeros(8)]
        self.lib.L.orc_sim_foot_coeffs(self.h, *[_p(x) for x in a])
        return a

    def zmp(self):
        zx, zy = np.zeros(100000), np.zeros(100000)
        n = self.lib.L.orc_sim_zmp(self.h, _p(zx), _p(zy), 100000)
        return zx[:n].copy(), zy[:n].copy()

    def frames(self):
        """Robot::getT / getX after the last updateState: (T [28][4][4], X [28][6][6])"""
        T, X = np.zeros((28, 4, 4)), np.zeros((28, 6, 6))
        self.lib.L.orc_sim_frames(self.h, _p(T), _p(X))
        return T, X

    def link_table(self):
        m, c, I = np.zeros(28), np.zeros((28, 3)), np.zeros((28, 3, 3))
        self.lib.L.orc_sim_link_table(self.h, _p(m), _p(c), _p(I))
        return m, c, I


def oracle():
    path = os.path.join(HERE, "liblmh_oracle.so")
    if not os.path.exists(path):
        build()
    return Lib(path)


def reference():
    """The reference's own sources (oracle/_ref). None if it was never built."""
    path = os.path.join(HERE, "_ref", "liblmh_ref.so")
    if not os.path.exists(path):
        try:
            build(ref=True)
        except Exception:
            return None
    return Lib(path) if os.path.exists(path) else None
