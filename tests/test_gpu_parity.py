"""GPU parity tests proper: the CUDA path, called through the C ABI (liblmh_b200.so), against the oracle on the
same seeded inputs, against the committed reference vectors, against 60-digit QP solutions, and -- at the full
BASELINE sizes -- through size-independent properties.  Tolerances are the north star's: single-step torques,
joint accelerations and contact forces 1e-6 relative; 1 s closed-loop CoM and foot trajectories 1e-4 m; integer
work (ZMP window index, step counters, push steps) bit-exact.  What we actually assert is much tighter."""
import numpy as np
import pytest

import linearmpchumanoid_b200 as lmh
from conftest import relerr, scalederr

pytestmark = pytest.mark.gpu

MODEL = ["M", "C", "Cg", "AG", "AGpqp", "Jpqp", "JFeet", "CoM", "comVel", "comAngMom", "xRef", "yRef", "qppRef",
         "hGpRef", "footAccRef"]


def _t34(T):
    return np.concatenate([T[:3, :3].ravel(), T[:3, 3]])


@pytest.fixture(scope="module", params=[32, 64, 128])
def nt(request):
    return request.param


@pytest.fixture(scope="module", params=["throughput", "latency"])
def kern(request):
    """the default (throughput) kernel: ten one-warp environments per block; and the opt-in latency kernel
    (envs_per_block = -1): one environment per block with two helper warps.  Both must pass every parity test."""
    return {} if request.param == "throughput" else {"envs_per_block": -1}


def test_native_library_is_the_thing_running():
    info = lmh.Batch(2).launch_info()
    assert info["threads_per_env"] in (32, 64, 128) and info["smem_bytes"] > 16000


def test_setup_matches_reference_vectors(golden):
    b = lmh.Batch(3)
    q0, zcom, mass, it = b.nominal()
    assert np.abs(q0 - golden["x0"][:30]).max() < 1e-13 and it == 4
    assert abs(zcom - float(golden["zcom"])) < 1e-14 and abs(mass - float(golden["mass"])) < 1e-14
    zx, zy, fc = b.plan(0)
    assert len(zx) == int(golden["zmp_len"]) and (zx == 0).all() and (zy == 0).all()
    assert np.abs(fc.ravel() - golden["foot_coef"]).max() < 1e-12
    x, vh, t = b.get_state()
    assert np.abs(x - np.concatenate([q0, np.zeros(30)])).max() == 0 and (vh == 0).all() and (t == 0).all()


def test_stage_against_reference_vectors(golden, nt):
    n = len(golden["stage_t"])
    b = lmh.Batch(n, threads_per_env=nt)
    b.set_state(golden["stage_x"], golden["stage_vh"], golden["stage_t"])
    d = b.stage()
    for i in range(n):
        for k in MODEL:
            assert scalederr(d[k][i], golden["stage_" + k][i]) < 1e-12, (i, k)
        assert scalederr(d["T7"][i], _t34(golden["stage_T7"][i])) < 1e-13
        assert relerr(d["Ha"][i], golden["stage_Ha"][i]) < 1e-12 and relerr(d["ga"][i], golden["stage_g"][i][:30]) < 1e-11
        for k in ("qpp", "tau", "f", "xdot"):
            assert relerr(d[k][i], golden["stage_" + k][i]) < 1e-8, (i, k)     # north star: 1e-6
        assert d["k_index"][i] == golden["stage_k_index"][i]                  # bit exact
    _, vh, _ = b.get_state()
    assert np.array_equal(vh, golden["stage_x"][:, 30:])                      # held velocity advanced like Robot::v_


def test_stage_against_oracle_random_states(oracle, nt):
    E = 96
    s = oracle.sim()
    x0, _, _ = s.get_state()
    rng = np.random.default_rng(1234)
    X, VH, T = np.tile(x0, (E, 1)), np.zeros((E, 30)), np.zeros(E)
    for e in range(E):
        sc = [0.0005, 0.002, 0.01, 0.04][e % 4]
        X[e, :30] += sc * rng.standard_normal(30); X[e, 30:] = 3 * sc * rng.standard_normal(30)
        VH[e] = X[e, 30:] + 0.2 * sc * rng.standard_normal(30); T[e] = 0.01 * (e % 40) + 0.005 * (e % 2)
    b = lmh.Batch(E, threads_per_env=nt)
    b.set_state(X, VH, T)
    d = b.stage()
    assert (d["qp_status"] == 0).all()
    for e in range(E):
        s.set_state(X[e], T[e], VH[e])
        do = s.stage(X[e], T[e])
        for k in MODEL:
            assert scalederr(d[k][e], do[k]) < 1e-12, (e, k)
        for k in ("qpp", "tau", "f"):
            assert relerr(d[k][e], do[k]) < 1e-7, (e, k)
        assert d["k_index"][e] == do["k_index"]
    assert d["qp_iters"].max() > 1                                            # the active-set path was exercised


def test_qp_against_exact_solutions(qp_exact):
    X = qp_exact["state"]
    b = lmh.Batch(len(X))
    b.set_state(X, X[:, 30:], np.zeros(len(X)))
    d = b.stage()
    assert (d["qp_status"] == 0).all() and (d["c"] >= 0).all()
    for i in range(len(X)):
        xe = qp_exact["x"][i]
        assert relerr(d["a"][i], xe[:30]) < 1e-8 and relerr(d["f"][i], xe[30:42]) < 1e-8, (i, int(qp_exact["n_active"][i]))
    assert int(qp_exact["n_active"].max()) >= 28                              # instances with nearly every bound active


def test_closed_loop_1s_against_reference_vectors(golden, nt):
    b = lmh.Batch(4, threads_per_env=nt, log_steps=100)          # (nt = 32, four environments: the helper-warp kernel)
    b.step(100)
    lg = b.log(0, 100)
    x, vh, t = b.get_state()
    for e in range(4):
        assert np.abs(lg["com_k4"][:, e] - golden["loop_com_k4"]).max() < 1e-10      # north star: 1e-4 m
        assert np.abs(lg["rfoot"][:, e] - golden["loop_rfoot"]).max() < 1e-10
        assert np.abs(lg["lfoot"][:, e] - golden["loop_lfoot"]).max() < 1e-10
        assert relerr(lg["tau_k1"][:, e], golden["loop_tau_k1"]) < 1e-8
        assert np.array_equal(lg["k_index"][:, e], golden["loop_k_index"])           # bit exact
        assert np.abs(x[e] - golden["loop_x"][-1]).max() < 1e-9
    assert t[0] == 100 * 0.01 or abs(t[0] - 1.0) < 1e-12
    tt = 0.0
    for _ in range(100):
        tt += 0.01
    assert t[0] == tt                                                                  # the clock accumulates like Clock::step


def test_stand_5s_stream(golden, kern):
    """config 1: the reference run, 500 steps; CoM x as printed by main.cpp:86"""
    b = lmh.Batch(1, log_steps=500, **kern)
    b.step(500)
    cx = b.log(0, 500)["com_k4"][:, 0, 0]
    assert np.abs(cx - golden["stand_comx"]).max() < 5e-8                              # stream has 6 printed digits
    st = b.stats()[0]
    assert st["steps"] == 500 and st["qp_fail"] == 0 and st["alive"] == 1


def test_pushed_rollout_against_oracle(oracle, kern):
    """config 2 in small: per-env pushes, 60 steps, every env against the CPU oracle"""
    E, S = 48, 60
    step, dv = lmh.push_plan(range(E), max_step=20)
    b = lmh.Batch(E, **kern)
    b.set_pushes(step, dv)
    b.step(25)
    b.step(S - 25)                                                                     # split launches: same result
    x, vh, t = b.get_state()
    xc, comc, stc, fails = oracle.batch_rollout(oracle.default_config(), E, S, push_step=step, push_dv=dv, n_threads=8)
    assert fails == 0
    assert np.abs(x - xc).max() < 1e-7
    assert np.abs(b.com() - comc).max() < 1e-9
    st = b.stats()
    assert relerr(st["max_abs_tau"], stc[:, 0]) < 1e-7
    assert (st["steps"] == S).all() and (st["alive"] == 1).all() and (st["qp_fail"] == 0).all()


def test_full_size_properties():
    """BASELINE size (4096 envs): properties that need no oracle"""
    E, S = 4096, 30
    step, dv = lmh.push_plan(range(E), max_step=10)
    b = lmh.Batch(E)
    b.set_pushes(step, dv)
    b.step(S)
    x, vh, t = b.get_state()
    st = b.stats()
    assert np.isfinite(x).all() and (st["alive"] == 1).all() and (st["qp_fail"] == 0).all() and (st["steps"] == S).all()
    # sharding invariance: any contiguous shard, run on its own, reproduces its slice bit for bit
    for lo, hi in ((0, 64), (1000, 1100), (4032, 4096)):
        b2 = lmh.Batch(hi - lo)
        b2.set_pushes(step[lo:hi], dv[lo:hi])
        b2.step(S)
        x2, _, _ = b2.get_state()
        assert np.array_equal(x2, x[lo:hi])
    # determinism: a second run of the same batch is bit-identical
    b.reset(); b.set_pushes(step, dv); b.step(S)
    assert np.array_equal(b.get_state()[0], x)
    # unpushed-yet environments are still identical copies of each other
    late = np.where(step >= S)[0]
    if len(late) > 1:
        assert np.array_equal(x[late[0]], x[late[1]])
    # mirror symmetry of the scenario: pushes (vx, vy) and (vx, -vy) at the same step give mirrored CoM motion
    b3 = lmh.Batch(2)
    b3.set_pushes(np.array([2, 2], np.int32), np.array([[0.1, 0.08, 0, 0, 0, 0], [0.1, -0.08, 0, 0, 0, 0]]))
    b3.step(40)
    c = b3.com()
    assert abs(c[0, 0] - c[1, 0]) < 1e-6 and abs(c[0, 1] + c[1, 1]) < 1e-6


def test_latency_and_throughput_kernels():
    """the throughput kernel gives the same bits whatever the block shape (1, 3 or 10 environments per block); the latency
    kernel (one environment per block with two helper warps: RNEA || CRBA || Jacobians) is deterministic and agrees with
    it to rounding -- it is another kernel, the compiler contracts a few sums of products differently"""
    E, S = 40, 50
    step, dv = lmh.push_plan(range(E), max_speed=0.5, max_step=20)
    out = []
    for kw in ({}, {"envs_per_block": 1}, {"envs_per_block": 3}, {"envs_per_block": -1}, {"envs_per_block": -1}):
        b = lmh.Batch(E, **kw)
        b.set_pushes(step, dv)
        b.step(S)
        x, vh, t = b.get_state()
        st = b.stats()
        out.append((x, vh, t, st["max_abs_tau"], st["qp_iters"], st["steps"], st["alive"]))
        b.close()
    for other in out[1:3]:
        for a, c in zip(out[0], other):
            assert np.array_equal(a, c)
    for a, c in zip(out[3], out[4]):
        assert np.array_equal(a, c)                                   # deterministic
    assert np.abs(out[3][0] - out[0][0]).max() < 1e-10 and np.array_equal(out[3][2], out[0][2])
    assert np.array_equal(out[3][4], out[0][4]) and np.array_equal(out[3][5], out[0][5])     # same iteration counts, same steps
    with pytest.raises(lmh.LmhError):
        lmh.Batch(4, envs_per_block=-2)


def test_thread_geometries_agree():
    xs = []
    for nt in (32, 64, 128):
        b = lmh.Batch(16, threads_per_env=nt)
        step, dv = lmh.push_plan(range(16), max_step=5)
        b.set_pushes(step, dv)
        b.step(20)
        xs.append(b.get_state()[0])
    # 64 and 128 threads run the same generic code; one warp per env runs the lane-role versions of the tree
    # recursions (lmh_warp.cuh): same results up to summation order
    assert np.array_equal(xs[1], xs[2]) and np.abs(xs[0] - xs[1]).max() < 1e-10


def test_plans_and_edge_cases():
    b = lmh.Batch(5)
    zx, zy, fc = b.plan(0)
    # two plans: the nominal one and a shifted ZMP reference; envs pick by id
    b.set_plans(np.stack([zx, zx + 0.01]), np.stack([zy, zy]), np.stack([fc, fc]), plan_ids=[0, 1, 0, 1, 1])
    d = b.stage()
    assert np.abs(d["xRef"][0] - d["xRef"][2]).max() == 0 and abs(d["xRef"][1][2] - d["xRef"][0][2]) > 1e-3
    with pytest.raises(lmh.LmhError):
        b.set_plans(zx[:10], zy[:10], fc)                                              # shorter than the MPC window
    with pytest.raises(lmh.LmhError):
        b.set_plans(np.stack([zx, zx]), np.stack([zy, zy]), np.stack([fc, fc]), plan_ids=[0, 1, 2, 0, 0])
    with pytest.raises(lmh.LmhError):
        lmh.Batch(0)
    with pytest.raises(lmh.LmhError):
        lmh.Batch(4, threads_per_env=48)
    b.step(0)                                                                          # empty rollout is a no-op
    assert (b.get_state()[2] == 0).all()
    # the ZMP window is clamped at the end of the reference array instead of reading past it
    b.set_state(t=np.full(5, 5.4))
    d = b.stage()
    assert (d["k_index"] == int(5.4 / 0.01)).all() and np.isfinite(d["xRef"]).all()
    # a hopeless state does not take the batch down: status is per env
    x, vh, t = b.get_state()
    x[3, :30] += 0.8
    b.set_state(x, vh, np.zeros(5))
    b.step(3)
    st = b.stats()
    assert np.isfinite(b.get_state()[0][[0, 1, 2, 4]]).all() and (st["steps"][[0, 1, 2, 4]] == 3).all()
    # ... and a fallen environment is frozen at its last state (finite), its statistics closed, alive = 0
    b.step(20)
    xs = b.get_state()[0]
    st = b.stats()
    assert np.isfinite(xs).all() and (st["alive"][[0, 1, 2, 4]] == 1).all()
    if st["alive"][3] == 0:
        x_frozen = xs[3].copy()
        b.step(5)
        assert np.array_equal(b.get_state()[0][3], x_frozen) and np.isfinite(b.stats()["max_abs_tau"]).all()


def test_walking_extension_against_oracle(oracle, kern):
    """EXTENSION (BASELINE configs 3-4): per-environment walking plans, 1.6 s MPC horizon, contact
    switching.  Defined in the oracle first; the GPU must agree env by env over 1.2 s of closed loop (north star: 1 s),
    and the contact-phase index / contact mask of every stage of every step must match as a SEQUENCE, bit exact."""
    E, n = 6, 120
    wp = lmh.walk_plan(range(E))
    b = lmh.Batch(E, time_horizon=1.6, log_steps=n, **kern)
    b.set_walk_plans(wp["zmp_x"], wp["zmp_y"], wp["phase_T"], wp["foot_coef"], wp["contact"], plan_ids=np.arange(E))
    b.step(n)
    lg = b.log(0, n)
    x, vh, t = b.get_state()
    for e in range(E):
        so = oracle.sim(time_horizon=1.6)
        so.set_walk_plan(wp["zmp_x"][e], wp["zmp_y"][e], wp["phase_T"], wp["foot_coef"][e], wp["contact"][e])
        lo = so.step(n)
        assert np.abs(lg["com_k4"][:, e] - lo["com_k4"]).max() < 1e-9, e      # north star: 1e-4 m
        assert np.abs(lg["rfoot"][:, e] - lo["rfoot"]).max() < 1e-9 and np.abs(lg["lfoot"][:, e] - lo["lfoot"]).max() < 1e-9
        assert np.abs(x[e] - lo["x"][-1]).max() < 1e-8
        assert relerr(lg["tau_k1"][:, e], lo["tau_k1"]) < 1e-7               # north star: 1e-6 relative
        assert np.array_equal(lg["k_index"][:, e], lo["k_index"])              # bit exact
        assert np.array_equal(lg["phase"][:, e], lo["phase"])                  # footstep / contact-phase index, whole sequence
        assert np.array_equal(lg["contact"][:, e], lo["contact"])
    assert lg["phase"].max() == 3 and set(np.unique(lg["contact"])) == {1, 2, 3}   # double support, both single supports
    st = b.stats()
    assert (st["qp_fail"] == 0).all() and (st["alive"] == 1).all()
    assert lg["rfoot"][:, :, 2].max() > 0.015 and lg["lfoot"][:, :, 2].max() > 0.015
    b.set_state(x, vh, np.full(E, 0.6))
    d = b.stage()
    assert (d["phase"] == 1).all() and np.abs(d["f"][:, :6]).max() == 0.0     # integer phase index; swing foot force-free
    with pytest.raises(lmh.LmhError):
        b.set_walk_plans(wp["zmp_x"], wp["zmp_y"], wp["phase_T"], wp["foot_coef"], wp["contact"] + 4, plan_ids=np.arange(E))


def test_config4_walk_and_push_against_oracle(oracle, kern):
    """BASELINE configs[3] in small: walking plans + pushes (seed 0x5EED0004) strong enough that some environments
    fall, 200 steps, every environment against the oracle.  Integer work -- alive, committed steps, QP failures, the
    contact-phase / contact-mask / ZMP-window sequences -- bit exact; survivors' states 1e-7, CoM and feet 1e-9 m;
    a fallen environment is closed at its last healthy step (finite state, clean statistics)."""
    E, S = 32, 200
    wp = lmh.walk_plan(range(E), seed=0x5EED0004)
    step, dv = lmh.push_plan(range(E), seed=0x5EED0004, max_speed=0.6, max_step=120)
    b = lmh.Batch(E, time_horizon=1.6, log_steps=S, **kern)
    b.set_walk_plans(wp["zmp_x"], wp["zmp_y"], wp["phase_T"], wp["foot_coef"], wp["contact"], plan_ids=np.arange(E))
    b.set_pushes(step, dv)
    b.step(90)
    b.step(S - 90)                                                             # split launches: fallen envs stay fallen
    x, vh, t = b.get_state()
    st, lg = b.stats(), b.log(0, S)
    r = oracle.batch_rollout_ex(oracle.default_config(time_horizon=1.6), E, S, push_step=step, push_dv=dv, walk=wp,
                                n_threads=8, log=True)
    so = r["stats"]
    assert 4 <= so["alive"].sum() <= E - 4                                     # a real mix of survivors and falls
    assert np.array_equal(st["alive"], so["alive"]) and np.array_equal(st["steps"], so["steps"])     # bit exact
    assert np.array_equal(st["qp_fail"], so["qp_fail"]) and np.array_equal(t, so["t"])
    assert np.isfinite(x).all() and (st["max_abs_tau"] < lmh.TAU_LIMIT).all() and (st["max_com_err"] < lmh.COM_ERR_LIMIT).all()
    for e in range(E):
        n, lo = int(st["steps"][e]), r["log"][e]
        for k in ("k_index", "phase", "contact"):
            assert np.array_equal(lg[k][:n, e], lo[k][:n]), (e, k)
        tol = 1e-9 if st["alive"][e] else 1e-5                                 # the last steps of a fall amplify round-off
        assert np.abs(lg["com_k4"][:n, e] - lo["com_k4"][:n]).max() < tol, e
        assert np.abs(lg["rfoot"][:n, e] - lo["rfoot"][:n]).max() < tol and np.abs(lg["lfoot"][:n, e] - lo["lfoot"][:n]).max() < tol
        assert abs(st["max_abs_tau"][e] - so["max_abs_tau"][e]) < 1e-4 * so["max_abs_tau"][e]
        assert abs(st["max_com_err"][e] - so["max_com_err"][e]) < 1e-6
        if st["alive"][e]:
            assert scalederr(x[e], r["x"][e]) < 1e-7, e                        # (a survivor may be tumbling at 40 rad/s)
            assert abs(st["max_abs_qpp"][e] - so["max_abs_qpp"][e]) < 1e-6 * so["max_abs_qpp"][e]
        else:
            assert np.abs(x[e] - r["x"][e]).max() < 1e-2, e
    # a fallen environment is not touched by further launches
    dead = np.where(st["alive"] == 0)[0]
    b.step(5)
    x2, _, t2 = b.get_state()
    assert np.array_equal(x2[dead], x[dead]) and np.array_equal(t2[dead], t[dead])
    assert np.array_equal(b.stats()["steps"][dead], st["steps"][dead])


def test_qp_microbenchmark_entry(oracle):
    """config 5: cold-started WBC QPs on frozen states through lmh_qp_bench, against the oracle's solver"""
    E = 32
    step, dv = lmh.push_plan(range(E), max_step=10)
    b = lmh.Batch(E)
    b.set_pushes(step, dv)
    b.step(15)
    X, VH, T = b.get_state()
    ms, a, f = b.qp_bench(0, 3)
    ms2, a2, f2 = b.qp_bench(0, 1)
    assert np.array_equal(a, a2) and np.array_equal(f, f2) and ms > 0        # repeats are cold starts: same answer
    so = oracle.sim()
    for e in range(E):
        so.set_state(X[e], T[e], VH[e])
        d = so.stage(X[e], T[e])
        assert relerr(a[e], d["x"][:30]) < 1e-7 and relerr(f[e], d["x"][30:42]) < 1e-7
    msm, _, _ = b.qp_bench(1, 5)
    assert msm > 0
    with pytest.raises(lmh.LmhError):
        b.qp_bench(2, 1)


def test_batched_inverse_kinematics(oracle, ik_golden):
    """section 8(f) rank 2: Kinematics::compute for a batch of targets on the device, against the reference build's
    postures, against the oracle on further seeded targets, and through the definition of the operational state"""
    g = ik_golden
    n = len(g["q"])
    b = lmh.Batch(n)
    r = b.ik(g["com"], g["rf"], g["lf"])
    assert (r["status"] == 0).all() and (r["residual"] <= 1e-10).all()
    assert np.abs(r["q"] - g["q"]).max() < 2e-8                      # both stop at |operational error| <= 1e-10; the
                                                                     # joint-space gap is that times the conditioning
    assert (r["iters"] <= g["iters"] + 1).all() and r["iters"][0] == 0   # starts from the nominal posture, not from scratch
    # further targets, live against the oracle; the nominal target must return the nominal posture at once
    from golden.make_ik_golden import targets
    com, rf, lf = targets(40, seed=7)
    ok = [i for i in range(40) if oracle.sim(com=com[i], rf=rf[i], lf=lf[i]).ik_iterations < 50]
    r2 = b.ik(com[ok], rf[ok], lf[ok])
    qo = np.stack([oracle.sim(com=com[i], rf=rf[i], lf=lf[i]).get_state()[0][:30] for i in ok])
    assert (r2["status"] == 0).all() and np.abs(r2["q"] - qo).max() < 2e-8
    q0 = b.nominal()[0]
    r3 = b.ik([[-0.02, 0.0, 0.26]], [0, -0.05, 0, 0, 0, 0], [0, 0.05, 0, 0, 0, 0])
    assert r3["iters"][0] == 0 and np.array_equal(r3["q"][0], q0)
    # an unreachable target ends at the iteration cap or a singular Jacobian instead of hanging (upstream's loop has no cap)
    r4 = b.ik([[-0.02, 0.0, 0.60]], [0, -0.05, 0, 0, 0, 0], [0, 0.05, 0, 0, 0, 0])
    assert r4["status"][0] in (1, 2)
    # apply: every environment restarts from its own posture; the stage evaluated there sees the targets
    r5 = b.ik(g["com"], g["rf"], g["lf"], apply=True)
    x, vh, t = b.get_state()
    assert np.array_equal(x[:, :30], r5["q"]) and (x[:, 30:] == 0).all() and (vh == 0).all() and (t == 0).all()
    d = b.stage()
    assert np.abs(d["CoM"] - g["com"]).max() < 2e-10
    assert np.abs(d["T7"][:, 9:] - g["rf"][:, :3]).max() < 2e-10 and np.abs(d["T14"][:, 9:] - g["lf"][:, :3]).max() < 2e-10
    # closed loop from those postures against the oracle started at the same states
    b.ik(g["com"], g["rf"], g["lf"], apply=True)
    b.step(20)
    xg = b.get_state()[0]
    xc, _, _, fails = oracle.batch_rollout(oracle.default_config(), n, 20, x0=np.hstack([r5["q"], np.zeros((n, 30))]), n_threads=8)
    assert fails == 0 and np.abs(xg - xc).max() < 1e-7


def test_placement_by_push_step_changes_nothing():
    """lmh_set_pushes sorts the SLOTS of the one-block-per-SM kernel by push step (environments that need extra QP
    iterations at the same time share a block); every environment must compute exactly what it computes unsorted
    (LmhConfig.placement = 1) and with fewer environments per block (LmhConfig.envs_per_block)"""
    E, S = 200, 40
    step, dv = lmh.push_plan(range(E), max_step=15)
    out = []
    for kw in (dict(), dict(placement=1), dict(envs_per_block=3)):
        b = lmh.Batch(E, **kw)
        b.set_pushes(step, dv)
        b.step(S)
        x, vh, t = b.get_state()
        st = b.stats()
        out.append((x, vh, t, st["max_abs_tau"], st["qp_iters"]))
        b.close()
    for other in out[1:]:
        for a, c in zip(out[0], other):
            assert np.array_equal(a, c)
    with pytest.raises(lmh.LmhError):
        lmh.Batch(4, envs_per_block=64)
    with pytest.raises(lmh.LmhError):
        b = lmh.Batch(4)
        b.set_pushes(step[:3], dv[:3])                                                # wrong-length arrays never reach C


def test_two_live_batches_with_different_horizons():
    """the dynamic shared-memory limit belongs to the process-wide kernel, not to a handle: a second batch with a shorter
    MPC horizon (smaller block-shared table) must not break the launches of the first"""
    b1 = lmh.Batch(40, time_horizon=1.6)
    b2 = lmh.Batch(40)                                                                 # 0.5 s horizon
    ref1, ref2 = lmh.Batch(40, time_horizon=1.6), None
    for _ in range(3):
        b1.step(2); b2.step(2)
    ref1.step(6)
    assert np.array_equal(b1.get_state()[0], ref1.get_state()[0])
    ref1.close()
    ref2 = lmh.Batch(40)
    ref2.step(6)
    assert np.array_equal(b2.get_state()[0], ref2.get_state()[0])


def test_setup_tables_through_the_c_abi(golden):
    """lmh_get_link_tables / lmh_get_mpc_gain against the reference vectors (Robot::getLinks after the constructor;
    r = beta Pu H^-1 e0 from the reference's own Px, Pu)"""
    b = lmh.Batch(2)
    m, c, isp = b.link_tables()
    assert np.array_equal(m, golden["link_mass"]) and np.abs(c - golden["link_com"]).max() < 1e-16
    assert abs(m.sum() - float(golden["mass"])) < 1e-14
    for i in range(28):                                       # spatial inertia blocks (Dynamics.cpp:4-13)
        if isp[i].any():
            assert np.allclose(isp[i][3:, 3:], m[i] * np.eye(3), atol=0) and np.abs(isp[i][:3, 3:] + isp[i][3:, :3]).max() == 0
            cx = isp[i][:3, 3:] / m[i]
            assert np.abs(np.array([cx[2, 1], cx[0, 2], cx[1, 0]]) - c[i]).max() < 1e-15
    r, rpx = b.mpc_gain()
    Pu, Px = golden["Pu"], golden["Px"]
    rr = Pu @ np.linalg.solve(1e-3 * np.eye(51) + Pu.T @ Pu, np.eye(51)[0])
    assert len(r) == 51 and np.abs(r - rr).max() < 1e-12 and np.abs(rpx - rr @ Px).max() < 1e-11


def test_grid_shape_changes_nothing():
    """whole waves of ten-environment blocks plus a thinned last wave (9 full blocks per SM would not divide 1483):
    an environment's result must not depend on how many neighbours its block has or where in the grid it sits"""
    S = 12
    ref = None
    for E in (1483, 37, 1):                      # 148 full blocks + 3 tail blocks of one; 37 blocks of one; one block
        b = lmh.Batch(E)
        step, dv = lmh.push_plan(range(E), max_step=6)
        b.set_pushes(step, dv)
        b.step(S)
        x, vh, t = b.get_state()
        info = b.launch_info()
        b.close()
        if ref is None:
            ref = (x, vh, t)
            assert info["blocks"] == 151
        else:
            assert np.array_equal(x, ref[0][:E]) and np.array_equal(vh, ref[1][:E]) and np.array_equal(t, ref[2][:E])


def test_kinematics_all_frames(golden, oracle):
    """Robot::updateState on the device for a batch of postures (lmh_kinematics): EVERY world transform T_i and
    parent-relative motion transform X_i (28 frames) against the reference vectors and, on further postures, the oracle"""
    q = golden["stage_x"][:, :30]
    b = lmh.Batch(2)
    T, X, com = b.kinematics(q)
    assert scalederr(T, golden["stage_T_all"]) < 1e-13 and scalederr(X, golden["stage_X_all"]) < 1e-13
    assert np.abs(com - golden["stage_CoM"]).max() < 1e-14
    assert (T[:, :, 3] == [0, 0, 0, 1]).all()
    rng = np.random.default_rng(5)
    q2 = golden["x0"][:30] + 0.3 * rng.standard_normal((7, 30))
    T2, X2, com2 = b.kinematics(q2)
    so = oracle.sim()
    for i in range(7):
        d = so.stage(np.concatenate([q2[i], np.zeros(30)]), 0.0)
        To, Xo = so.frames()
        assert scalederr(T2[i], To) < 1e-13 and scalederr(X2[i], Xo) < 1e-13 and np.abs(com2[i] - d["CoM"]).max() < 1e-14


def test_mpc_dense_solver_on_the_device(oracle):
    """N2: Mpc3dLip::compute as a DENSE solve per problem (gradient by a dense product, shared Cholesky factor, forward and
    back substitution), FP64 FMA and FP64 tensor-core (DMMA) builds, against numpy's dense solve of the reference's own
    formulation (Px, Pu from Mpc3dLip::initialize), against the oracle's Mpc3dLip, and against the gain-vector form the
    closed-loop kernel uses.  Horizons of BASELINE configs[4] (16, 32, 64) plus the app's 50 and the walking workloads' 160."""
    rng = np.random.default_rng(11)
    for N in (16, 32, 50, 64, 160):                    # 160 = the walking workloads' 1.6 s horizon
        hz = N * 0.01 * (1.0 + 1e-9)                      # (int)(horizon / dt) == N as mpcLinearPendulum.hpp:49 computes it
        b = lmh.Batch(2, time_horizon=hz)
        Px, Pu = b.mpc_tables()
        assert Px.shape == (N + 1, 2)
        n = 37                                            # not a multiple of the 16 / 32 problems a block takes
        com, vel = 0.05 * rng.standard_normal((n, 2)), 0.2 * rng.standard_normal((n, 2))
        zx, zy = 0.03 * rng.standard_normal((n, N + 1)), 0.03 * rng.standard_normal((n, N + 1))
        H = 1e-3 * np.eye(N + 1) + Pu.T @ Pu
        A, B = np.array([[1, 0.01], [0, 1]]), np.array([0.01 ** 2 / 2, 0.01])
        ref = np.zeros((2, n, 3))
        for ax, z in enumerate((zx, zy)):
            for i in range(n):
                xk = np.array([com[i, ax], vel[i, ax]])
                u = np.linalg.solve(H, -(Pu.T @ (Px @ xk - z[i])))
                ref[ax, i] = np.concatenate([A @ xk + B * u[0], [u[0]]])
        for dmma in (False, True):
            xr, yr, ms = b.mpc_solve(com, vel, zx, zy, use_dmma=dmma)
            assert np.abs(xr - ref[0]).max() < 1e-10 and np.abs(yr - ref[1]).max() < 1e-10, (N, dmma)
            xr2, yr2, _ = b.mpc_solve(com, vel, zx, zy, use_dmma=dmma, repeats=3)     # repeats recompute from scratch
            assert np.array_equal(xr, xr2) and np.array_equal(yr, yr2) and ms > 0
        # the oracle's Mpc3dLip (dense factorisation at this horizon) through a stage evaluation at constant ZMP reference
        so = oracle.sim(time_horizon=hz)
        x0 = so.get_state()[0]
        d = so.stage(x0, 0.0)
        c0 = d["CoM"][:2]
        xr, yr, _ = b.mpc_solve([c0], [[0.0, 0.0]], np.zeros((1, N + 1)), np.zeros((1, N + 1)), use_dmma=True)
        assert np.abs(xr[0] - d["xRef"]).max() < 1e-10 and np.abs(yr[0] - d["yRef"]).max() < 1e-10
        dg = b.stage()                                    # gain-vector form inside the closed-loop kernel
        assert np.abs(dg["xRef"][0] - xr[0]).max() < 1e-10 and np.abs(dg["yRef"][0] - yr[0]).max() < 1e-10
        b.close()
    with pytest.raises(lmh.LmhError):
        lmh.Batch(1).mpc_solve([[0, 0]], [[0, 0]], np.zeros((1, 7)), np.zeros((1, 7)))


def test_general_inverse_kinematics_entry(ik_golden):
    """lmh_ik_solve: full operational targets and per-target starting postures (Kinematics::desiredOperationalState +
    compute as the class surface calls them)"""
    g = ik_golden
    n = len(g["q"])
    b = lmh.Batch(1)
    q0 = b.nominal()[0]
    des = np.zeros((n, 30))
    des[:, 0:6], des[:, 6:12], des[:, 12:24], des[:, 27:30] = g["rf"], g["lf"], q0[18:30], g["com"]
    r = b.ik_solve(des)
    assert (r["status"] == 0).all() and np.abs(r["q"] - g["q"]).max() < 2e-8
    qi = np.array([-0.0185, 0, 0.282, 0, 0, 0, 0, 0, -0.5, 0.8, -0.3, 0, 0, 0, -0.5, 0.8, -0.3, 0, 1.6, 0, 0, 0, 0, -1.6, 0, 0, 0, 0, 0, 0])
    r2 = b.ik_solve(des[:3], q_start=qi)                 # from Robot's initialConfiguration (Robot.cpp:242-251), as upstream does
    assert (r2["status"] == 0).all() and np.abs(r2["q"] - g["q"][:3]).max() < 2e-8 and (r2["iters"] > 0).all()
    des2 = des[:2].copy(); des2[:, 12] += 0.2             # a different arm target is honoured
    r3 = b.ik_solve(des2)
    assert (r3["status"] == 0).all() and np.abs(r3["q"][:, 18] - des2[:, 12]).max() < 1e-10


def test_walk_plans_generated_inside_the_library(oracle):
    """lmh_make_walk_plans (ZMP::walkZMP + per-phase foot plans behind the C ABI, keyed by the global environment id)
    installs exactly what lmh_walk_plan returns: same rollouts bit for bit, for a shard that does not start at id 0"""
    E, lo, S = 24, 1000, 60
    wp = lmh.walk_plan(range(lo, lo + E), seed=0x5EED0003)
    a = lmh.Batch(E, time_horizon=1.6)
    a.set_walk_plans(wp["zmp_x"], wp["zmp_y"], wp["phase_T"], wp["foot_coef"], wp["contact"], plan_ids=np.arange(E))
    b = lmh.Batch(E, time_horizon=1.6)
    b.make_walk_plans(first_env_id=lo, seed=0x5EED0003)
    a.step(S); b.step(S)
    assert np.array_equal(a.get_state()[0], b.get_state()[0])
    assert not np.array_equal(a.get_state()[0][0], a.get_state()[0][1])          # the plans differ per environment
    ws = lmh.walk_plan([7], fixed_len=0.035)                                       # ZMP(Task::Walk, ...): one nominal step length
    assert ws["step_len"][0] == 0.035 and ws["zmp_x"][0].max() > 0.06


def test_full_size_walking_sweep_properties(oracle):
    """BASELINE sizes of configs 3-4 (16384 and 32768 environments per GPU): properties that need no oracle at that size --
    clean per-environment statistics, determinism, sharding invariance of a shard that does not start at 0 -- plus the
    oracle on a handful of environments picked out of the big batch"""
    E, S, lo = 32768, 30, 5000
    step, dv = lmh.push_plan(range(E), seed=0x5EED0004, max_speed=0.6, max_step=20)
    b = lmh.Batch(E, time_horizon=1.6)
    b.make_walk_plans(first_env_id=0, seed=0x5EED0004)
    b.set_pushes(step, dv)
    b.step(S)
    x, vh, t = b.get_state()
    st = b.stats()
    assert np.isfinite(x).all() and (st["qp_fail"] == 0).all()
    assert (st["max_abs_tau"] < lmh.TAU_LIMIT).all() and (st["max_com_err"] < lmh.COM_ERR_LIMIT).all()   # no garbage maxima
    assert ((st["steps"] == S) == (st["alive"] == 1)).all() and st["alive"].mean() > 0.9
    b.close()
    # a shard of 16384 environments starting at global id 5000 reproduces its slice bit for bit
    b2 = lmh.Batch(16384, time_horizon=1.6)
    b2.make_walk_plans(first_env_id=lo, seed=0x5EED0004)
    b2.set_pushes(step[lo:lo + 16384], dv[lo:lo + 16384])
    b2.step(S)
    x2 = b2.get_state()[0]
    st2 = b2.stats()
    assert np.array_equal(x2, x[lo:lo + 16384]) and np.array_equal(st2["steps"], st["steps"][lo:lo + 16384])
    b2.close()
    # ... and eight environments of the big batch against the oracle
    ids = [0, 4097, 9999, 16383, 16384, 20000, 31111, 32767]
    wp = lmh.walk_plan(ids, seed=0x5EED0004)
    r = oracle.batch_rollout_ex(oracle.default_config(time_horizon=1.6), len(ids), S, push_step=step[ids], push_dv=dv[ids], walk=wp,
                                n_threads=8)
    assert np.array_equal(r["stats"]["alive"], st["alive"][ids]) and np.array_equal(r["stats"]["steps"], st["steps"][ids])
    ok = st["alive"][ids] == 1
    assert scalederr(x[ids][ok], r["x"][ok]) < 1e-8
