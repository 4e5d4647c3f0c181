"""The oracle against (1) vectors generated from the reference's own sources (tests/golden/reference_vectors.npz),
(2) the reference build itself when it is present, (3) first-principles known answers,
(4) 60-digit QP solutions.  CPU only."""
import numpy as np
import pytest

from conftest import relerr

STAGE_KEYS = ["M", "C", "Cg", "AG", "AGpqp", "Jpqp", "JFeet", "CoM", "comVel", "comAngMom", "xRef", "yRef", "qppRef",
              "hGpRef", "footAccRef", "g", "qpp", "tau", "f", "T7", "T14", "xdot"]


def test_setup_matches_reference_vectors(oracle, golden):
    s = oracle.sim()
    x0, _, _ = s.get_state()
    assert np.array_equal(x0, golden["x0"])                      # IK posture, bit for bit
    assert s.mass == float(golden["mass"]) and s.zcom == float(golden["zcom"])
    Px, Pu, hz = s.mpc_tables()
    assert hz == int(golden["horizon"]) == 50                    # App. B.6
    assert np.array_equal(Px, golden["Px"]) and np.array_equal(Pu, golden["Pu"])
    assert np.array_equal(np.concatenate(s.foot_coeffs()), golden["foot_coef"])
    zx, zy = s.zmp()
    assert len(zx) == int(golden["zmp_len"]) == 550              # (int)((T+0.5)/dt), App. B.4
    m, c, I = s.link_table()
    assert np.array_equal(m, golden["link_mass"]) and np.array_equal(c, golden["link_com"])
    assert np.array_equal(I, golden["link_inertia"])


def test_stage_matches_reference_vectors(oracle, golden):
    s = oracle.sim(second_wbc=1, mpc_rebuild_hessian=1)          # the configuration the reference always runs
    for i in range(len(golden["stage_t"])):
        x, vh, t = golden["stage_x"][i], golden["stage_vh"][i], float(golden["stage_t"][i])
        s.set_state(x, t, vh)
        d = s.stage(x, t)
        for k in STAGE_KEYS:
            assert np.array_equal(d[k], golden["stage_" + k][i]), (i, k)
        assert d["k_index"] == int(golden["stage_k_index"][i])
        assert np.array_equal(d["x"][:30], golden["stage_a"][i])


def test_closed_loop_matches_reference_vectors(oracle, golden):
    s = oracle.sim(second_wbc=1, mpc_rebuild_hessian=1)
    lg = s.step(100)
    for k in ("x", "com_k4", "rfoot", "lfoot", "tau_k1"):
        assert np.array_equal(lg[k], golden["loop_" + k]), k
    assert np.array_equal(lg["k_index"], golden["loop_k_index"])  # integer work: bit exact


def test_cheaper_oracle_modes_change_nothing_that_matters(oracle, golden):
    """bench baseline mode: one WBC solve per stage and a cached MPC factor."""
    s = oracle.sim(second_wbc=0, mpc_rebuild_hessian=0)
    lg = s.step(100)
    assert np.abs(lg["x"] - golden["loop_x"]).max() < 1e-12
    assert np.abs(lg["tau_k1"] - golden["loop_tau_k1"]).max() < 1e-9
    assert np.array_equal(lg["k_index"], golden["loop_k_index"])


def test_stand_stream(oracle, golden):
    """the reference's only observable output: CoM x per step (main.cpp:86), 500 steps for T = 5 (App. B.4)"""
    s = oracle.sim()
    cx = s.run_stand()
    ref = golden["stand_comx"]
    assert len(cx) == len(ref) == 500
    assert np.abs(cx - ref).max() < 5e-8                          # the stream is printed with 6 significant digits
    assert abs(cx[0] + 0.02) < 1e-4 and abs(cx[-1]) < 2e-4        # converges from -0.02 towards the ZMP reference 0
    assert (np.diff(cx[:50]) > 0).all() and np.abs(cx[100:]).max() < 1e-3


def test_against_reference_build_live(oracle, reflib):
    if reflib is None:
        pytest.skip("oracle/_ref not built and /root/reference absent")
    so, sr = oracle.sim(second_wbc=1, mpc_rebuild_hessian=1), reflib.sim()
    rng = np.random.default_rng(5)
    x0, _, _ = sr.get_state()
    for i in range(4):
        x = x0.copy(); x[:30] += 0.01 * rng.standard_normal(30); x[30:] = 0.05 * rng.standard_normal(30)
        vh = 0.05 * rng.standard_normal(30)
        so.set_state(x, 0.02 * i, vh); sr.set_state(x, 0.02 * i, vh)
        do, dr = so.stage(x, 0.02 * i), sr.stage(x, 0.02 * i)
        for k in STAGE_KEYS + ["H", "x"]:
            assert np.array_equal(do[k], dr[k]), k
    lo, lr = oracle.sim(second_wbc=1, mpc_rebuild_hessian=1).step(30), reflib.sim().step(30)
    assert np.array_equal(lo["x"], lr["x"]) and np.array_equal(lo["k_index"], lr["k_index"])


# ---- first-principles known answers (the reference ships no tests of its own)
def test_physics_identities(oracle):
    s = oracle.sim()
    x0, _, _ = s.get_state()
    assert abs(s.mass - 5.30539) < 1e-12                          # sum of the 28 mass literals
    rng = np.random.default_rng(3)
    x = x0.copy(); x[:30] += 0.02 * rng.standard_normal(30); x[30:] = 0.2 * rng.standard_normal(30)
    s.set_state(x, 0.0, x[30:])
    d = s.stage(x, 0.0)
    M = d["M"]
    assert np.abs(M[6:, 6:] - M[6:, 6:].T).max() < 1e-15 and np.abs(M[:6, 6:] - M[6:, :6].T).max() == 0
    assert np.linalg.eigvalsh(0.5 * (M + M.T)).min() > 0
    # mass block: m*I up to the reference's non-orthonormal hip pre-rotations (0.7071, App. B.1): 2*0.7071^2 = 1 - 1.6e-5
    assert np.abs(M[3:6, 3:6] - s.mass * np.eye(3)).max() < 5e-4 and abs(M[3, 3] - s.mass) < 1e-12
    # RNEA is linear in the base acceleration: C - Cg = M[:, 3:6] R0' g
    eta = x[3:6]
    cr, sr_, cp, sp, cy, sy = np.cos(eta[0]), np.sin(eta[0]), np.cos(eta[1]), np.sin(eta[1]), np.cos(eta[2]), np.sin(eta[2])
    R0 = np.array([[cy * cp, cy * sp * sr_ - sy * cr, cy * sp * cr + sy * sr_], [sy * cp, sy * sp * sr_ + cy * cr, sy * sp * cr - cy * sr_], [-sp, cp * sr_, cp * cr]])
    assert np.abs(d["C"] - d["Cg"] - M[:, 3:6] @ (R0.T @ [0, 0, 9.81])).max() < 1e-12
    # contact forces: c >= 0, f = G c rows satisfied, dynamics rows satisfied, torque identity
    xq = d["x"]; a, f = xq[:30], xq[30:42]
    assert xq[42:].min() > -1e-7 and d["kkt"][1] < 1e-7
    assert np.abs(M[:6] @ a - d["JFeet"].T[:6] @ f + d["C"][:6]).max() < 1e-8
    assert np.abs((M @ a + d["C"] - d["JFeet"].T @ f)[6:] - d["tau"]).max() < 1e-12
    # vertical force balance at rest: sum f_z = m (g + zdd_com)
    s2 = oracle.sim(); d0 = s2.stage(x0, 0.0)
    com_acc_z = (d0["AG"] @ d0["x"][:30] + d0["AGpqp"])[5] / s.mass
    # (holds to ~1e-4 relative only: the reference's leg transforms are not exactly rigid, App. B.1)
    assert abs(d0["f"][5] + d0["f"][11] - s.mass * (9.81 + com_acc_z)) < 1e-2


def test_jacobian_and_momentum_by_finite_differences(oracle):
    s = oracle.sim()
    x0, _, _ = s.get_state()
    rng = np.random.default_rng(11)
    v = 0.3 * rng.standard_normal(30)
    x = x0.copy(); x[30:] = v
    s.set_state(x, 0.0, v)
    d = s.stage(x, 0.0)
    h = 1e-6
    qd = d["xdot"][:30]                                           # generalized coordinate rates of this velocity
    xp, xm = x.copy(), x.copy(); xp[:30] += h * qd; xm[:30] -= h * qd
    dp, dm = oracle.sim().stage(xp, 0.0), oracle.sim().stage(xm, 0.0)
    # agreement is limited to ~1e-4 relative by the reference's non-rigid leg transforms (App. B.1), not by FD error
    assert np.abs((dp["CoM"] - dm["CoM"]) / (2 * h) - d["comVel"]).max() < 1e-4
    for T, rows in (("T7", slice(3, 6)), ("T14", slice(9, 12))):
        vfd = (dp[T][:3, 3] - dm[T][:3, 3]) / (2 * h)
        # JFeet acts on the base-frame twist form of v; comVel already uses it, so take J v from footAccRef terms:
        Jv = d["JFeet"][rows] @ _base_twist(x, v)
        assert np.abs(vfd - Jv).max() < 1e-4


def _base_twist(x, v):
    eta, p = x[3:6], x[0:3]
    cr, sr_, cp, sp, cy, sy = np.cos(eta[0]), np.sin(eta[0]), np.cos(eta[1]), np.sin(eta[1]), np.cos(eta[2]), np.sin(eta[2])
    R0 = np.array([[cy * cp, cy * sp * sr_ - sy * cr, cy * sp * cr + sy * sr_], [sy * cp, sy * sp * sr_ + cy * cr, sy * sp * cr - cy * sr_], [-sp, cp * sr_, cp * cr]])
    w, vl = v[3:6], v[0:3]
    out = v.copy()
    out[0:3] = R0.T @ w
    out[3:6] = R0.T @ (vl + np.cross(w, p))
    return out


def test_mpc_gain_known_answer(oracle):
    """u0 = -r'(Px xk - z); for z = 0 this is a state feedback on [pos, vel]"""
    s = oracle.sim()
    Px, Pu, N = s.mpc_tables()
    H = 1e-3 * np.eye(N + 1) + Pu.T @ Pu
    r = Pu @ np.linalg.solve(H, np.eye(N + 1)[0])
    gain = -(r @ Px)
    assert np.allclose(gain, [-20.855, -7.707], atol=2e-3)
    x0, _, _ = s.get_state()
    d = s.stage(x0, 0.0)
    assert abs(d["xRef"][2] - gain @ [d["CoM"][0], d["comVel"][0]]) < 1e-10


def test_time_index_sequence(oracle):
    """k = (int)(t/dt) with t accumulated by t += dt differs from the step counter (F10): replay it."""
    s = oracle.sim()
    lg = s.step(500)
    t, dt, ks = 0.0, 0.01, []
    for _ in range(500):
        ks.append([int(t / dt), int((t + 0.5 * dt) / dt), int((t + 0.5 * dt) / dt), int((t + dt) / dt)])
        t += dt
    ks = np.array(ks)
    assert np.array_equal(lg["k_index"], ks)
    assert (ks[:, 0] != np.arange(500)).sum() > 100              # the trap is real


def test_qp_solver_against_exact_solutions(oracle, qp_exact):
    worst = 0.0
    for i in range(len(qp_exact["g"])):
        H, g, A, lb = qp_exact["H"][i], qp_exact["g"][i], qp_exact["A"][i], qp_exact["lbA"][i]
        ub = lb.copy(); ub[18:] = 1e20
        x, y, st, it, kkt = oracle.qp_solve(H, g, A, lb, ub)
        assert st == 0
        xe = qp_exact["x"][i]
        worst = max(worst, relerr(x[:30], xe[:30]), relerr(x[30:42], xe[30:42]))
    assert worst < 1e-7, worst                                    # (a, f) are what parity is judged on; c is eps-regularised


def test_inverse_kinematics_matches_reference_vectors(oracle, ik_golden):
    """Kinematics::compute on seeded targets: restatement vs the reference build's postures (bit for bit when made)"""
    g = ik_golden
    for i in range(len(g["q"])):
        s = oracle.sim(com=g["com"][i], rf=g["rf"][i], lf=g["lf"][i])
        assert np.abs(s.get_state()[0][:30] - g["q"][i]).max() < 1e-12 and s.ik_iterations == g["iters"][i]


def test_health_rule_definition(oracle):
    """EXTENSION (orc_capi.h): a step is committed only if healthy; a fallen environment is frozen at its last healthy state
    with its clock and statistics closed, and a healthy one is untouched by the rule"""
    cfg = oracle.default_config()
    E, S = 3, 80
    step = np.array([5, 5, -1], np.int32)
    dv = np.zeros((E, 6)); dv[0, 0] = 2.5; dv[1, 0] = 0.05                      # env 0 is thrown over, env 1 nudged, env 2 left alone
    r = oracle.batch_rollout_ex(cfg, E, S, push_step=step, push_dv=dv, n_threads=3, log=True)
    st = r["stats"]
    assert list(st["alive"]) == [0, 1, 1] and st["steps"][0] < S and list(st["steps"][1:]) == [S, S]
    n0 = int(st["steps"][0])
    tt = 0.0
    for _ in range(n0):
        tt += 0.01
    assert st["t"][0] == tt and np.isfinite(r["x"][0]).all()                     # clock stopped with the environment
    assert np.array_equal(r["log"][0]["x"][n0 - 1], r["x"][0])                   # frozen at the last committed state
    assert np.array_equal(r["log"][0]["x"][S - 1], r["x"][0])                    # ... and still there at the end
    assert st["max_abs_tau"][0] < 100.0 and st["max_com_err"][0] < 0.5          # the failing step never reached the statistics
    # a healthy environment equals the plain rollout bit for bit (the rule only decides, it does not compute)
    x2, _, _, _ = oracle.batch_rollout(cfg, 1, S)
    assert np.array_equal(r["x"][2], x2[0])
    assert (r["log"][2]["phase"] == 0).all() and (r["log"][2]["contact"] == 3).all()      # stand task: one phase, both feet
