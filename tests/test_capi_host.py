"""Host-side checks of the drop-in boundary that need no GPU: the shared library loads, exports every
symbol include/lmh_capi.h declares, its struct layouts match the Python binding, and it refuses to run
without a CUDA device instead of falling back to anything."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import linearmpchumanoid_b200 as lmh
from linearmpchumanoid_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    return open(os.path.join(ROOT, "include", "lmh_capi.h")).read()


def test_library_is_built_in_tree():
    assert os.path.exists(lmh.LIB_PATH), "run __graft_entry__.build() first"
    assert os.path.dirname(lmh.LIB_PATH) == os.path.join(ROOT, "linearmpchumanoid_b200")


def test_exports_every_declared_symbol():
    declared = sorted(set(re.findall(r"\b(lmh_[a-z0-9_]+)\s*\(", _header())))
    assert declared == sorted(api.SYMBOLS)
    L = lmh.load_library()
    for sym in declared:
        assert hasattr(L, sym), sym


def test_struct_layouts_match_header():
    h = _header()
    # LmhStageDump: count the doubles declared in the header
    body = h[h.index("typedef struct {\n    double M[900]"):h.index("} LmhStageDump;")]
    n_dbl = sum(int(n) for n in re.findall(r"\w+\[(\d+)\]", body.split("int32_t")[0]))
    assert api.STAGE_DTYPE.itemsize == 8 * n_dbl + 16
    assert api.STEPLOG_DTYPE.itemsize == 8 * (3 + 3 + 3 + 24) + 4 * 14
    assert api.STATS_DTYPE.itemsize == 8 * 3 + 4 * 6
    assert C.sizeof(api.LmhConfig) == 8 * (3 + 3 + 6 + 6 + 3) + 4 * 6 + 8


def test_default_config_is_the_reference_scenario():
    L = lmh.load_library()
    cfg = api.LmhConfig()
    L.lmh_default_config(C.byref(cfg))
    assert (cfg.dt, cfg.sim_time, cfg.time_horizon) == (0.01, 5.0, 0.5)          # main.cpp:13-14,38
    assert list(cfg.com) == [-0.02, 0.0, 0.26]                                   # main.cpp:31
    assert list(cfg.rf)[1] == -0.05 and list(cfg.lf)[1] == 0.05                  # main.cpp:24-27
    assert cfg.mu == 0.7 and cfg.mpc_alpha == 1e-3 and cfg.mpc_beta == 1.0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is only reachable without one")
    with pytest.raises(lmh.LmhError, match="no CUDA device"):
        lmh.Batch(4)
    with pytest.raises(lmh.LmhError):
        lmh.measure_fp64_peak(0)


def test_product_does_not_touch_the_oracle():
    """the product package and its native sources never import, include or link anything under oracle/"""
    pkg = os.path.join(ROOT, "linearmpchumanoid_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".inc")):
                txt = open(os.path.join(dp, f)).read()
                assert "orc_" not in txt and "oracle/" not in txt.replace("# oracle", ""), f
    out = os.popen("ldd " + lmh.LIB_PATH).read()
    assert "oracle" not in out and "hostemu" not in out


def test_mujoco_export_glue():
    """relabelMujocoMatrix (apps/mujoco/main.cpp:182-200) and the free-joint pose of MujocoSim.cpp:77-85, written out
    by hand: head, left leg, right leg, left arm (shoulder pitch negated), right arm; host-only data-format glue"""
    q = np.arange(60, dtype=float).reshape(2, 30) / 7.0
    o = lmh.export_mujoco(q)
    L = np.zeros((24, 24))
    L[0:2, 22:24] = np.eye(2); L[2:8, 6:12] = np.eye(6); L[8:14, 0:6] = np.eye(6)
    L[14:19, 17:22] = np.eye(5); L[14, 17] = -1; L[19:24, 12:17] = np.eye(5)
    for e in range(2):
        assert np.array_equal(o[e, 7:], L @ q[e, 6:])
        assert np.array_equal(o[e, :7], [q[e, 0], q[e, 1], q[e, 2] + 0.035, 1, 0, 0, 0])
