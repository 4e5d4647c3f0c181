#!/usr/bin/env python3
"""Generates tests/golden/reference_vectors.npz from the REFERENCE'S OWN SOURCES
(oracle/_ref/liblmh_ref.so = /root/reference/src/*.cpp + apps/offline/main.cpp compiled against the shim
Eigen/qpOASES headers of oracle/refbuild).  Run in the build container only (needs /root/reference):

    make -C oracle ref && python tests/golden/make_golden.py

What the vectors pin: everything the reference computes per stage and in closed loop, except the two
third-party boundaries (Eigen's summation order, qpOASES' solve), which the shims stand in for.
"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import orc_py

R = orc_py.reference()
assert R is not None, "oracle/_ref/liblmh_ref.so missing: run `make -C oracle ref`"
s = R.sim()
x0, _, _ = s.get_state()
out = {"x0": x0, "mass": s.mass, "zcom": s.zcom}
Px, Pu, hz = s.mpc_tables()
out.update(Px=Px, Pu=Pu, horizon=hz)
fc = s.foot_coeffs()
out["foot_coef"] = np.concatenate(fc)
zx, zy = s.zmp()
out.update(zmp_len=len(zx), zmp_x_unique=np.unique(zx), zmp_y_unique=np.unique(zy))
m, c, I = s.link_table()
out.update(link_mass=m, link_com=c, link_inertia=I)

# stage dumps at seeded states (state, held velocity, time) -> every intermediate
rng = np.random.default_rng(20261018)
KEYS = ["M", "C", "Cg", "AG", "AGpqp", "Jpqp", "JFeet", "CoM", "comVel", "comAngMom", "xRef", "yRef", "qppRef",
        "hGpRef", "footAccRef", "g", "qpp", "tau", "f", "T7", "T14", "xdot"]
X, VH, T, dumps = [], [], [], {k: [] for k in KEYS + ["Ha", "a", "k_index", "qp_iters", "T_all", "X_all"]}
for i in range(12):
    x = x0.copy(); vh = np.zeros(30); t = 0.0
    if i:
        sc = [0.001, 0.003, 0.01][i % 3]
        x[:30] += sc * rng.standard_normal(30); x[30:] = 3 * sc * rng.standard_normal(30)
        vh = x[30:] + 0.1 * sc * rng.standard_normal(30); t = 0.01 * i + 0.005 * (i % 2)
    s.set_state(x, t, vh)
    d = s.stage(x, t)
    X.append(x); VH.append(vh); T.append(t)
    for k in KEYS: dumps[k].append(d[k])
    dumps["Ha"].append(d["H"][:30, :30]); dumps["a"].append(d["x"][:30])
    dumps["k_index"].append(d["k_index"]); dumps["qp_iters"].append(d["qp_iters"])
    Tall, Xall = s.frames()                       # Robot::getT / getX: every world transform and parent-relative motion transform
    dumps["T_all"].append(Tall); dumps["X_all"].append(Xall)
out.update(stage_x=np.array(X), stage_vh=np.array(VH), stage_t=np.array(T))
for k, v in dumps.items(): out["stage_" + k] = np.array(v)

# closed loop: 100 steps (1 s) with per-step logs, then the whole 5 s CoM-x stream of stand() (main.cpp:86)
s2 = R.sim()
lg = s2.step(100)
for k in ("x", "com_k4", "rfoot", "lfoot", "tau_k1", "k_index"): out["loop_" + k] = lg[k]
s3 = R.sim()
out["stand_comx"] = s3.run_stand()
np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
print("wrote reference_vectors.npz:", {k: np.shape(v) for k, v in out.items()})
