"""The device source (lmh_model.cuh + lmh_setup.hpp) compiled for the host with NT = 1 (tests/hostemu, a test
tool) against the oracle and the reference vectors: kernel LOGIC parity without a GPU.  The GPU parity tests
proper are in test_gpu_parity.py."""
import numpy as np

from conftest import relerr, scalederr

MODEL = ["M", "C", "Cg", "AG", "AGpqp", "Jpqp", "JFeet", "CoM", "comVel", "comAngMom", "xRef", "yRef", "qppRef",
         "hGpRef", "footAccRef", "T7", "T14"]


def _t34(T):          # oracle 4x4 -> (R row-major 9, p 3)
    return np.concatenate([T[:3, :3].ravel(), T[:3, 3]])


def test_setup_tables(hostemu, oracle, golden):
    e = hostemu.Emu()
    q0, zcom, mass, it = e.nominal()
    assert np.abs(q0 - golden["x0"][:30]).max() < 1e-13 and it == 4          # Newton IK, same root
    assert abs(zcom - float(golden["zcom"])) < 1e-14 and abs(mass - float(golden["mass"])) < 1e-14
    tb = e.tables()
    assert np.abs(tb["mass"] - golden["link_mass"]).max() == 0
    assert np.abs(tb["com"] - golden["link_com"]).max() < 1e-16
    assert np.abs(tb["foot_coef"] - golden["foot_coef"]).max() < 1e-12
    # MPC gain vector against its definition r = beta Pu H^-1 e0 from the reference's Px/Pu
    Pu, Px = golden["Pu"], golden["Px"]
    H = 1e-3 * np.eye(51) + Pu.T @ Pu
    r = Pu @ np.linalg.solve(H, np.eye(51)[0])
    assert np.abs(tb["mpc_r"] - r).max() < 1e-12 and np.abs(tb["rPx"] - r @ Px).max() < 1e-11


def test_stage_against_reference_vectors(hostemu, golden):
    e = hostemu.Emu()
    for i in range(len(golden["stage_t"])):
        d, vh = e.stage(golden["stage_x"][i], golden["stage_vh"][i], float(golden["stage_t"][i]))
        for k in MODEL:
            ref = golden["stage_" + k][i]
            if k in ("T7", "T14"):
                ref = _t34(ref)
            assert scalederr(d[k], ref) < 1e-12, (i, k)
        assert relerr(d["Ha"], golden["stage_Ha"][i]) < 1e-12 and relerr(d["ga"], golden["stage_g"][i][:30]) < 1e-11
        for k in ("qpp", "tau", "f", "xdot"):                                # north-star tolerance: 1e-6 relative
            assert relerr(d[k], golden["stage_" + k][i]) < 1e-8, (i, k)
        assert d["k_index"] == int(golden["stage_k_index"][i])               # integer work: bit exact
        assert np.array_equal(vh, golden["stage_x"][i][30:])                 # held velocity advanced


def test_qp_against_exact_solutions(hostemu, qp_exact):
    """(a, f) of the reduced primal active-set solve vs 60-digit certified minimisers, incl. 20-31 active bounds"""
    e = hostemu.Emu()
    worst = 0.0
    for i in range(len(qp_exact["g"])):
        x = qp_exact["state"][i]
        d, _ = e.stage(x, x[30:], 0.0)
        assert d["qp_status"] == 0
        xe = qp_exact["x"][i]
        worst = max(worst, relerr(d["a"], xe[:30]), relerr(d["f"], xe[30:42]))
        assert d["c"].min() >= 0.0                                           # feasibility is exact by construction
        assert ((d["c"] == 0) == (xe[42:] <= 1e-14)).all() or int(qp_exact["n_active"][i]) > 0
    assert worst < 1e-8, worst


def test_qp_working_set_table(hostemu):
    """Rows of the working-set table against the problem they tabulate: for the free set F of one foot and a target g in
    wrench space, c_F = argmin 1/2 wF |G_F c|^2 + 1/2 eps |c|^2 - g'G_F c  and  f = G_F c_F.  Checked in extended precision
    for the full set, flat sub-patterns, an edge, a single vertex, single generators, the empty set and random subsets."""
    e = hostemu.Emu()
    rng = np.random.default_rng(5)
    masks = [0xFFFF, 0x0000, 0x00FF, 0xFF00, 0x0F0F, 0x000F, 0xF000, 0x0001, 0x8000, 0x0101, 0x7FFF, 0xFFFE, 0x1248]
    masks += [int(m) for m in rng.integers(1, 0xFFFF, 40)]
    ld = np.longdouble
    for m in masks:
        Kf, Yf, Gf, wF, eps = e.qp_table_row(m)
        free = [g for g in range(16) if (m >> g) & 1]
        if not free:
            assert not Kf.any() and not Yf.any()
            continue
        G = Gf[:, free].astype(ld)
        H = ld(wF) * (G.T @ G) + ld(eps) * np.eye(len(free), dtype=ld)
        # H^-1 G' by Gaussian elimination in extended precision (numpy.linalg has no longdouble)
        A = np.concatenate([H, G.T], axis=1)
        n = len(free)
        for k in range(n):
            piv = k + int(np.argmax(np.abs(A[k:, k])))
            A[[k, piv]] = A[[piv, k]]
            A[k] = A[k] / A[k, k]
            for i in range(n):
                if i != k:
                    A[i] = A[i] - A[i, k] * A[k]
        C = A[:, n:]                                    # c_F = C g
        Kx = (G @ C).astype(np.float64)
        scale = np.abs(Kx).max()
        assert np.abs(Kf - Kx).max() < 1e-10 * scale, hex(m)
        assert np.abs(Kf - Kf.T).max() == 0.0
        cF = (Gf[:, free].T @ Yf)                       # c_F = G_F' Yf g for every g
        assert np.abs(cF - C.astype(np.float64)).max() < 1e-10 * np.abs(C.astype(np.float64)).max(), hex(m)


def test_closed_loop_1s(hostemu, golden):
    e = hostemu.Emu()
    x, vh, t, lg, st = e.step(golden["x0"], np.zeros(30), 0.0, 100)
    assert np.abs(lg["com_k4"] - golden["loop_com_k4"]).max() < 1e-10       # north star: 1e-4 m
    assert np.abs(lg["rfoot"] - golden["loop_rfoot"]).max() < 1e-10
    assert np.abs(lg["lfoot"] - golden["loop_lfoot"]).max() < 1e-10
    assert np.abs(x - golden["loop_x"][-1]).max() < 1e-9
    assert relerr(lg["tau_k1"], golden["loop_tau_k1"]) < 1e-8
    assert np.array_equal(lg["k_index"], golden["loop_k_index"])           # bit exact
    assert st["steps"] == 100 and st["qp_fail"] == 0 and st["alive"] == 1


def test_random_states_against_oracle(hostemu, oracle):
    e = hostemu.Emu()
    s = oracle.sim()
    x0, _, _ = s.get_state()
    rng = np.random.default_rng(99)
    for i in range(20):
        sc = [0.001, 0.005, 0.02, 0.08][i % 4]
        x = x0.copy(); x[:30] += sc * rng.standard_normal(30); x[30:] = 3 * sc * rng.standard_normal(30)
        vh = x[30:] + 0.2 * sc * rng.standard_normal(30); t = 0.01 * i
        s.set_state(x, t, vh)
        do = s.stage(x, t)
        d, _ = e.stage(x, vh, t)
        for k in MODEL[:-2]:
            assert scalederr(d[k], do[k]) < 1e-12, (i, k)
        for k in ("qpp", "tau", "f"):
            assert relerr(d[k], do[k]) < 1e-7, (i, k, sc)


def test_walking_extension_against_oracle(hostemu, oracle):
    """EXTENSION (configs 3-4): phase plans with contact switching; kernel logic vs the oracle's definition."""
    import linearmpchumanoid_b200 as lmh
    wp = lmh.walk_plan(range(3, 4))
    so = oracle.sim(time_horizon=1.6)
    so.set_walk_plan(wp["zmp_x"][0], wp["zmp_y"][0], wp["phase_T"], wp["foot_coef"][0], wp["contact"][0])
    e = hostemu.Emu(horizon=1.6)
    e.set_walk_plan(wp["zmp_x"][0], wp["zmp_y"][0], wp["phase_T"], wp["foot_coef"][0], wp["contact"][0])
    x0, _, _ = so.get_state()
    n = 90                                                       # through the first swing phase and into the second
    lo = so.step(n)
    x, vh, t, lg, st = e.step(x0, np.zeros(30), 0.0, n)
    assert np.abs(lg["com_k4"] - lo["com_k4"]).max() < 1e-9 and np.abs(lg["rfoot"] - lo["rfoot"]).max() < 1e-9
    assert np.abs(x - lo["x"][-1]).max() < 1e-8 and relerr(lg["tau_k1"], lo["tau_k1"]) < 1e-7
    assert np.array_equal(lg["k_index"], lo["k_index"])
    # contact-phase index and contact mask of every stage of every step: the whole sequence, bit exact
    assert np.array_equal(lg["phase"], lo["phase"]) and np.array_equal(lg["contact"], lo["contact"])
    assert lg["phase"].max() == 2 and set(np.unique(lg["contact"])) == {1, 2, 3}
    assert st["qp_fail"] == 0 and st["alive"] == 1
    assert lo["rfoot"][:, 2].max() > 0.015                       # the right foot really left the ground ...
    d, _ = e.stage(lo["x"][59], lo["x"][59][30:], 0.6)          # ... and carries no force while it swings (phase 1)
    assert d["phase"] == 1 and np.abs(d["f"][:6]).max() == 0.0 and d["f"][11] > 40.0


def test_config4_walk_and_push_with_falls(hostemu, oracle):
    """config 4 in small (kernel logic on the host): walking plans + pushes large enough that environments fall.  The
    health rule (include/lmh_capi.h) must close a fallen environment at the same step as the oracle: alive, steps and
    qp_fail bit exact, statistics of the committed steps equal, no garbage maxima."""
    import linearmpchumanoid_b200 as lmh
    ids, S = [1, 4, 8, 10, 15], 200
    wp = lmh.walk_plan(ids, seed=0x5EED0004)
    step, dv = lmh.push_plan(ids, seed=0x5EED0004, max_speed=0.6, max_step=120)
    r = oracle.batch_rollout_ex(oracle.default_config(time_horizon=1.6), len(ids), S, push_step=step, push_dv=dv, walk=wp,
                                n_threads=5, log=True)
    so = r["stats"]
    x0 = oracle.sim(time_horizon=1.6).get_state()[0]
    assert 0 < so["alive"].sum() < len(ids)                      # some fall, some do not
    for i in range(len(ids)):
        e = hostemu.Emu(horizon=1.6)
        e.set_walk_plan(wp["zmp_x"][i], wp["zmp_y"][i], wp["phase_T"], wp["foot_coef"][i], wp["contact"][i])
        x, vh, t, lg, st = e.step(x0, np.zeros(30), 0.0, S, push_step=int(step[i]), push_dv=dv[i])
        n = int(st["steps"])
        assert (n, int(st["alive"]), int(st["qp_fail"])) == (so["steps"][i], so["alive"][i], so["qp_fail"][i])   # bit exact
        assert t == so["t"][i]                                   # the clock of a fallen environment stops with it
        lo = r["log"][i]
        for k in ("k_index", "phase", "contact"):
            assert np.array_equal(lg[k][:n], lo[k][:n]), k
        tol = 1e-9 if st["alive"] else 1e-5                      # the last steps of a fall amplify round-off
        assert np.abs(lg["com_k4"][:n] - lo["com_k4"][:n]).max() < tol
        assert np.abs(lg["rfoot"][:n] - lo["rfoot"][:n]).max() < tol and np.abs(lg["lfoot"][:n] - lo["lfoot"][:n]).max() < tol
        assert st["max_abs_tau"] < 100.0 and abs(st["max_abs_tau"] - so["max_abs_tau"][i]) < 1e-4 * so["max_abs_tau"][i]
        assert abs(st["max_com_err"] - so["max_com_err"][i]) < 1e-6
        if st["alive"]:
            assert scalederr(x, r["x"][i]) < 1e-7
        else:                                                    # frozen at the last healthy state, finite
            assert np.isfinite(x).all() and np.abs(x - r["x"][i]).max() < 1e-2
