#!/usr/bin/env python3
"""Generates tests/golden/qp_exact.npz: WBC QP instances (controller.cpp:388-479 data H, g, A, lbA, ubA in the
reference's own 74-variable / 50-row form) at states that activate contact-coefficient bounds, together with
their minimisers computed in 60-digit arithmetic (mpmath) and CERTIFIED by the exact KKT conditions
(primal feasibility, multipliers of active bounds >= 0).  The active set comes from a float64 primal
active-set run; the certificate is independent of how it was found.  These pin the QP boundary that the
reference delegates to qpOASES (absent from this image, version unpinned upstream).

    python tests/golden/make_qp_exact.py        (build container; takes about a minute)
"""
import os, sys
import numpy as np
import mpmath as mp
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import orc_py
mp.mp.dps = 60
O = orc_py.oracle()
s = O.sim()
x0, _, _ = s.get_state()

mu = 0.7
Bf = np.array([[mu, 0, 1], [0, mu, 1], [-mu, 0, 1], [0, -mu, 1]]).T
verts = np.array([[0.1, 0.025, 0], [0.1, -0.025, 0], [-0.05, 0.025, 0], [-0.05, -0.025, 0]])
def cross(p): return np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]])
Gf = np.zeros((6, 16))
for v in range(4):
    Gf[0:3, 4 * v:4 * v + 4] = cross(verts[v]) @ Bf
    Gf[3:6, 4 * v:4 * v + 4] = Bf
G = np.zeros((12, 32)); G[0:6, 0:16] = Gf; G[6:12, 16:32] = Gf

def qp_data(d):
    """the reference's A (50x74), lbA: rows 0-5 dynamics, 6-17 friction equalities, 18-49 c >= 0"""
    A = np.zeros((50, 74)); lb = np.zeros(50)
    A[0:6, 0:30] = d["M"][0:6, :]; A[0:6, 30:42] = -d["JFeet"].T[0:6, :]; lb[0:6] = -d["C"][0:6]
    for foot in range(2):
        n0, f0, c0 = 30 + 6 * foot, 33 + 6 * foot, 42 + 16 * foot
        rf, rn = 6 + 6 * foot, 9 + 6 * foot
        A[rf:rf + 3, f0:f0 + 3] = -np.eye(3); A[rn:rn + 3, n0:n0 + 3] = -np.eye(3)
        A[rf:rf + 3, c0:c0 + 16] = Gf[3:6]; A[rn:rn + 3, c0:c0 + 16] = Gf[0:3]
    A[18:50, 42:74] = np.eye(32)
    return A, lb

def active_set_f64(H, g, A, lb):
    """float64 primal active set on the reduced problem (variables a, c); returns the free mask of c."""
    Ha, ga = H[:30, :30], g[:30]; M6, Jb = A[0:6, 0:30], -A[0:6, 30:42].T; C6 = -lb[0:6]
    La = np.linalg.cholesky(Ha); Y = np.linalg.solve(La, M6.T); w = np.linalg.solve(La, ga)
    Sa = Y.T @ Y; rhs = C6 - Y.T @ w; Ls = np.linalg.cholesky(Sa)
    Wm = np.linalg.solve(Ls, Jb.T @ G); Q = G.T @ G + Wm.T @ Wm + 1e-8 * np.eye(32); p = Wm.T @ np.linalg.solve(Ls, rhs)
    def eqp(F):
        idx = np.where(F)[0]; cs = np.zeros(32)
        if len(idx): cs[idx] = np.linalg.solve(Q[np.ix_(idx, idx)], p[idx])
        return cs, Q @ cs - p
    F = np.ones(32, bool); cs, m_ = eqp(F)
    if (cs >= 0).all(): return F
    c = np.maximum(cs, 0); F = cs > 0
    for _ in range(200):
        cs, m_ = eqp(F)
        if (cs[F] >= 0).all():
            c = cs; viol = np.where(~F, -m_, -np.inf); i = viol.argmax()
            if viol[i] <= 1e-12 * max(1.0, np.abs(p).max()): return F
            F[i] = True
        else:
            dl = cs - c; r = np.where(F & (cs < 0), c / np.maximum(-dl, 1e-300), np.inf); i = r.argmin()
            c = c + min(1.0, r[i]) * dl; c[i] = 0; F[i] = False
    raise RuntimeError("no convergence")

def exact(H, g, A, lb, F):
    rows = list(range(18)) + [18 + i for i in range(32) if not F[i]]
    Aa = mp.matrix(A[rows].tolist()); b = mp.matrix(lb[rows].tolist()); m = len(rows)
    K = mp.zeros(74 + m, 74 + m); K[0:74, 0:74] = mp.matrix(H.tolist()); K[0:74, 74:] = Aa.T; K[74:, 0:74] = Aa
    sol = mp.lu_solve(K, mp.matrix([-v for v in g.tolist()] + list(b)))
    x = np.array([float(sol[i]) for i in range(74)])
    y = np.zeros(50)          # qpOASES sign: H x + g = A' y, y >= 0 at an active lower bound
    for k, r in enumerate(rows): y[r] = float(-sol[74 + k])
    c = x[42:]
    ok = bool((c[F] >= 0).all() and (y[18:][~F] >= -1e-30).all())
    return x, y, ok

rng = np.random.default_rng(7)
Hs, gs, As, lbs, xs, ys, nact, states = [], [], [], [], [], [], [], []
trial = 0
while len(xs) < 24:
    sc = [0.0, 0.002, 0.01, 0.03, 0.1, 0.3][trial % 6]; trial += 1
    x = x0.copy(); x[:30] += 0.3 * sc * rng.standard_normal(30); x[30:] = 3 * sc * rng.standard_normal(30)
    s.set_state(x, 0.0, x[30:]); d = s.stage(x, 0.0)
    H = 0.5 * (d["H"] + d["H"].T); g = d["g"]; A, lb = qp_data(d)
    F = active_set_f64(H, g, A, lb)
    xe, ye, ok = exact(H, g, A, lb, F)
    if not ok:
        print("trial", trial, "not certified, skipped"); continue
    Hs.append(H); gs.append(g); As.append(A); lbs.append(lb); xs.append(xe); ys.append(ye); nact.append(int((~F).sum())); states.append(x)
    print("instance", len(xs), "scale", sc, "active bounds", nact[-1], "|a|max %.2e" % np.abs(xe[:30]).max())
np.savez_compressed(os.path.join(HERE, "qp_exact.npz"), H=np.array(Hs), g=np.array(gs), A=np.array(As), lbA=np.array(lbs),
                    x=np.array(xs), y=np.array(ys), n_active=np.array(nact), state=np.array(states))
print("wrote qp_exact.npz")
