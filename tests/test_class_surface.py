"""The drop-in boundary: the reference's C++ class surface (include/linearMpcHumanoid/**: Robot, Dynamics, Kinematics,
Mpc3dLip, ZMP, Controller, footCoeffTrajectory, rk4Step, the generalizedFunctions helpers) on top of the C ABI.

CPU part (no GPU): the reference's OWN apps/offline/main.cpp compiles and links UNCHANGED against the include tree (where
/root/reference exists), so do our app and the probe program, and without a CUDA device they fail loudly.
GPU part: the prebuilt reference app (oracle/_ref/offline_main_on_b200, built by `make -C oracle ref_app` from the
unmodified upstream file) prints the reference's 500-line CoM-x stream; the probe program's numbers match the
reference vectors."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import relerr, scalederr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "linearmpchumanoid_b200")
REF_MAIN = "/root/reference/apps/offline/main.cpp"
REF_APP = os.path.join(ROOT, "oracle", "_ref", "offline_main_on_b200")


def _compile(src, exe, extra=()):
    cmd = ["g++", "-O1", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"), src, "-L" + LIBDIR, "-llmh_b200",
           "-Wl,-rpath," + LIBDIR, "-o", exe, *extra]
    return subprocess.run(cmd, capture_output=True, text=True)


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("probe") / "surface_probe")
    r = _compile(os.path.join(ROOT, "tests", "cpp", "surface_probe.cpp"), exe)
    assert r.returncode == 0, r.stderr
    return exe


def test_reference_main_compiles_unchanged(tmp_path):
    if not os.path.exists(REF_MAIN):
        pytest.skip("/root/reference is only present in the build container")
    exe = str(tmp_path / "offline_reference")
    r = _compile(REF_MAIN, exe)
    assert r.returncode == 0, r.stderr          # the unmodified upstream file against include/linearMpcHumanoid/**
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref_app"], check=True, capture_output=True)
    assert os.path.exists(REF_APP)


def test_our_app_and_probe_compile_and_refuse_to_run_without_a_gpu(tmp_path, probe):
    exe = str(tmp_path / "offline_b200")
    r = _compile(os.path.join(ROOT, "apps", "offline_b200", "main.cpp"), exe)
    assert r.returncode == 0, r.stderr
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is only reachable without one")
    for e in (exe, probe):
        out = subprocess.run([e], capture_output=True, text=True)
        assert out.returncode != 0 and "no CUDA device" in out.stderr          # loud, not a fallback


def test_host_side_types_and_helpers(tmp_path):
    """CPU only: the Eigen-API subset the headers fall back to without Eigen3 (comma initialiser, segment / head / tail views,
    the aliasing assignments of main.cpp:111-116, the arithmetic of rk4Step) and the generalizedFunctions helpers against
    their definitions; findPolyCoeff goes through the library's host-only entry point"""
    exe = str(tmp_path / "eigen_lite_probe")
    r = _compile(os.path.join(ROOT, "tests", "cpp", "eigen_lite_probe.cpp"), exe, extra=("-Wextra",))
    assert r.returncode == 0 and "warning" not in r.stderr, r.stderr
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout + out.stderr


def test_class_surface_has_the_reference_headers():
    want = ["robotInfo/Robot.hpp", "robotInfo/linkInertia.hpp", "robotInfo/robotParameters.hpp", "controller/controller.hpp",
            "controller/Dynamics.hpp", "controller/invKinematics.hpp", "controller/mpcLinearPendulum.hpp",
            "trajectories/zmpGeneration.hpp", "trajectories/footRefTrajectory.hpp", "general/Clock.hpp", "general/Task.hpp",
            "general/rk4.hpp", "general/generalizedFunctions.hpp"]
    for w in want:
        assert os.path.exists(os.path.join(ROOT, "include", "linearMpcHumanoid", w)), w


def _parse(text):
    d = {}
    for line in text.strip().split("\n"):
        p = line.split()
        d[p[0]] = np.array([float(v) for v in p[1:]])
    return d


@pytest.mark.gpu
def test_reference_app_prints_the_reference_stream(golden):
    """apps/offline/main.cpp of the reference, unmodified, on this library: the 500 CoM-x values of main.cpp:86"""
    if not os.path.exists(REF_APP):
        pytest.skip("oracle/_ref/offline_main_on_b200 was not built (needs /root/reference at build time)")
    out = subprocess.run([REF_APP], check=True, capture_output=True, text=True).stdout.split()
    cx = np.array([float(v) for v in out])
    assert len(cx) == 500
    assert np.abs(cx - golden["stand_comx"]).max() < 5e-8          # both streams carry 6 significant digits


@pytest.mark.gpu
def test_our_app_stand_fused_and_walk(golden, tmp_path):
    exe = str(tmp_path / "offline_b200")
    assert _compile(os.path.join(ROOT, "apps", "offline_b200", "main.cpp"), exe).returncode == 0
    cx = np.array([float(v) for v in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()])
    assert len(cx) == 500 and np.abs(cx - golden["stand_comx"]).max() < 5e-8
    fused = subprocess.run([exe, "--fused", "64"], check=True, capture_output=True, text=True).stdout.split()
    assert abs(float(fused[0]) - golden["stand_comx"][-1]) < 5e-8   # the fused batched rollout ends in the same place
    wk = np.array([float(v) for v in subprocess.run([exe, "--walk"], check=True, capture_output=True, text=True).stdout.split()])
    assert len(wk) == 240 and wk[-1] > 0.05                          # four 3.5 cm steps: the CoM has moved forward
    wf = subprocess.run([exe, "--walk", "--fused", "8"], check=True, capture_output=True, text=True).stdout.split()
    assert float(wf[0]) > 0.04


@pytest.mark.gpu
def test_class_surface_numbers(golden, probe):
    p = _parse(subprocess.run([probe], check=True, capture_output=True, text=True).stdout)
    g = golden
    assert abs(p["mass"][0] - float(g["mass"])) < 1e-14 and list(p["frames"]) == [28, 30, 24, 25]
    assert np.array_equal(p["link_mass"], g["link_mass"]) and np.abs(p["link_com3"] - g["link_com"][3]).max() < 1e-16
    assert np.abs(p["link_inertia3"].reshape(3, 3) - g["link_inertia"][3]).max() < 1e-18
    assert p["ik_iterations"][0] == 4 and np.abs(p["q0"] - g["x0"][:30]).max() < 1e-9
    assert np.abs(p["com0"] - g["stage_CoM"][0]).max() < 1e-10
    assert scalederr(p["T7"].reshape(4, 4), g["stage_T_all"][0][7]) < 1e-9 and scalederr(p["T27"].reshape(4, 4), g["stage_T_all"][0][27]) < 1e-9
    assert scalederr(p["X5"].reshape(6, 6), g["stage_X_all"][0][5]) < 1e-9
    assert list(p["tree"]) == [13, 12, 4]
    # Dynamics / Kinematics getters at the IK posture (the posture itself agrees to 1e-9, so do the terms)
    for k, shape in (("M", (30, 30)), ("C", 30), ("Cg", 30), ("AG", (6, 30)), ("AGpqp", 6), ("Jpqp", 12), ("JFeet", (12, 30))):
        assert scalederr(p[k].reshape(shape), g["stage_" + k][0]) < 1e-8, k
    assert np.array_equal(p["JR"], p["JFeet"][:180]) and np.array_equal(p["JL"], p["JFeet"][180:])
    # Robot::updateVelocityState / computeComMomentum (Robot.cpp:271-274,300-310) at v = 0.01 everywhere, from the reference's
    # own AG and X_0: swap [v_lin, w] -> [w, v_lin], express in the base frame, h = AG v
    AG, X0, v = g["stage_AG"][0], g["stage_X_all"][0][0], np.full(30, 0.01)
    vb = v.copy(); vb[:6] = X0 @ np.concatenate([v[3:6], v[0:3]])
    h = AG @ vb
    assert np.abs(p["comVel"] - h[3:] / float(g["mass"])).max() < 1e-9 and np.abs(p["angMom"] - h[:3]).max() < 1e-9
    assert list(p["zmp_sizes"][:2]) == [550, 150] and p["zmp_sizes"][2] == 0.05
    assert np.abs(p["xRef"] - g["stage_xRef"][0]).max() < 1e-9 and np.abs(p["yRef"] - g["stage_yRef"][0]).max() < 1e-9
    assert p["yRef_left"][2] != 0.0 and abs(p["zcom"][0] - float(g["zcom"])) < 1e-9 and p["zcom"][1] == 0.26
    fc = g["foot_coef"]
    assert np.abs(np.concatenate([p["rF0"], p["rF1"], p["rF2"]]) - fc[:20]).max() < 1e-12 and np.abs(p["lF1"] - fc[26:32]).max() < 1e-12
    for k in ("qpp", "tau", "f"):
        assert relerr(p[k], g["stage_" + k][0]) < 1e-6, k                  # north star: 1e-6 relative (posture from our own IK)
    assert np.array_equal(p["tau"], p["tau_get"]) and np.abs(p["mpc_xRef_after"] - g["stage_xRef"][0]).max() < 1e-9
    # foot coefficients handed to the Controller reach the QP
    assert np.abs(p["footAccRef_moving"] - p["footAccRef_still"]).max() > 0.1
    assert np.abs(p["footAccRef_moving"][6:] - p["footAccRef_still"][6:]).max() < 1e-9      # the left foot was not asked to move
    assert abs(p["rMove0"][0]) < 1e-15 and abs(np.polyval(p["rMove0"][::-1], 0.4) - 0.03) < 1e-12
    assert abs(np.polyval(p["rMove2"][::-1], 0.2) - 0.02) < 1e-12
    # walking through the class surface: phase 1 at t = 0.6 s, the swinging right foot carries no force
    assert p["walk"][1] == 6 and p["walk"][2] > 0.0 and abs(p["walk"][3] - 0.05) < 1e-15
    assert p["walk_stage"][0] == 1 and np.abs(p["walk_stage"][1:7]).max() == 0.0 and p["walk_stage"][12] > 40.0
    assert p["caught"][0] == 3
