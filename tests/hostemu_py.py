"""ctypes view of tests/hostemu/libhostemu.so (TEST TOOL: the device source compiled for the host)."""
import ctypes as C
import os

import numpy as np

from linearmpchumanoid_b200.api import STAGE_DTYPE, STEPLOG_DTYPE, STATS_DTYPE

_L = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostemu", "libhostemu.so"))
_f64, _i32, _vp = C.c_double, C.c_int32, C.c_void_p
_L.emu_create.restype = _vp
_L.emu_create.argtypes = [_f64, _f64, _f64, _vp]
_L.emu_destroy.argtypes = [_vp]
_L.emu_nominal.argtypes = [_vp] * 5
_L.emu_tables.argtypes = [_vp] * 8
_L.emu_stage.argtypes = [_vp, _vp, _vp, _f64, _vp]
_L.emu_qp_table_row.argtypes = [_vp, _i32, _vp, _vp, _vp]
_L.emu_set_walk_plan.argtypes = [_vp, _i32, _vp, _vp, _i32, _f64, _vp, _vp]
_L.emu_step.argtypes = [_vp, _vp, _vp, _vp, _i32, _vp, _vp]
_L.emu_rollout.argtypes = [_vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Emu:
    def __init__(self, dt=0.01, sim_time=5.0, horizon=0.5):
        self.h = _L.emu_create(dt, sim_time, horizon, None)

    def __del__(self):
        if getattr(self, "h", None):
            _L.emu_destroy(self.h)
            self.h = None

    def nominal(self):
        q, z, m, it = np.zeros(30), np.zeros(1), np.zeros(1), np.zeros(1, np.int32)
        _L.emu_nominal(self.h, _p(q), _p(z), _p(m), _p(it))
        return q, float(z[0]), float(m[0]), int(it[0])

    def tables(self):
        m, c, i, r, n1, rpx, fc = (np.zeros(28), np.zeros((28, 3)), np.zeros((28, 6, 6)), np.zeros(512),
                                   np.zeros(1, np.int32), np.zeros(2), np.zeros(40))
        _L.emu_tables(self.h, _p(m), _p(c), _p(i), _p(r), _p(n1), _p(rpx), _p(fc))
        return dict(mass=m, com=c, isp=i, mpc_r=r[:int(n1[0])].copy(), rPx=rpx, foot_coef=fc)

    def set_walk_plan(self, zx, zy, phase_T, foot_coef, contact):
        zx = np.ascontiguousarray(zx, np.float64); zy = np.ascontiguousarray(zy, np.float64)
        fc = np.ascontiguousarray(foot_coef, np.float64); ct = np.ascontiguousarray(contact, np.int32)
        _L.emu_set_walk_plan(self.h, len(zx), _p(zx), _p(zy), len(ct), float(phase_T), _p(fc), _p(ct))

    def qp_table_row(self, mask):
        """(Kf [6][6], Yf [6][6], Gf [6][16], wForce, epsC) of the working set `mask` (bit g = generator g free)"""
        row, gf, we = np.zeros(72), np.zeros(96), np.zeros(2)
        _L.emu_qp_table_row(self.h, int(mask), _p(row), _p(gf), _p(we))
        return row[:36].reshape(6, 6), row[36:].reshape(6, 6), gf.reshape(6, 16), float(we[0]), float(we[1])

    def stage(self, x, v_held, t):
        d = np.zeros(1, STAGE_DTYPE)
        vh = np.ascontiguousarray(v_held, np.float64).copy()
        _L.emu_stage(self.h, _p(np.ascontiguousarray(x, np.float64)), _p(vh), float(t), _p(d))
        return d[0], vh

    def step(self, x, v_held, t, n, push_step=-1, push_dv=None):
        x = np.ascontiguousarray(x, np.float64).copy()
        vh = np.ascontiguousarray(v_held, np.float64).copy()
        tt = np.array([float(t)])
        lg, st = np.zeros(n, STEPLOG_DTYPE), np.zeros(1, STATS_DTYPE)
        dv = None if push_dv is None else np.ascontiguousarray(push_dv, np.float64)
        _L.emu_rollout(self.h, _p(x), _p(vh), _p(tt), n, int(push_step), _p(dv), _p(lg), _p(st))
        return x, vh, float(tt[0]), lg, st[0]
