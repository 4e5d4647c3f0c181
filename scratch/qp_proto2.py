"""c-space primal active set: Q = G'KG + eps I (32x32), K = wF I + Jb Sa^-1 Jb'."""
import sys; sys.path.insert(0,'/root/repo/scratch')
from qp_proto import G, eps, wF, O
import numpy as np

def solve_c(d, maxit=100):
    Ha = d['H'][:30,:30]; ga = d['g'][:30]
    M6 = d['M'][:6,:]; C6 = d['C'][:6]; Jb = d['JFeet'][:,:6]
    La = np.linalg.cholesky(Ha)
    Y = np.linalg.solve(La, M6.T); w = np.linalg.solve(La, ga)
    Sa = Y.T@Y; rhs = C6 - Y.T@w
    Ls = np.linalg.cholesky(Sa)
    Wm = np.linalg.solve(Ls, Jb.T@G)          # 6x32
    Q = wF*G.T@G + Wm.T@Wm + eps*np.eye(32)
    p = Wm.T@np.linalg.solve(Ls, rhs)          # G'Jb Sa^-1 rhs
    def eqp(F):
        idx = np.where(F)[0]; cs = np.zeros(32)
        if len(idx):
            L = np.linalg.cholesky(Q[np.ix_(idx,idx)])
            cs[idx] = np.linalg.solve(L.T, np.linalg.solve(L, p[idx]))
        return cs, Q@cs - p      # gradient = multipliers on W
    F = np.ones(32,bool); cs, mu = eqp(F); it=1
    if (cs>=0).all(): c = cs
    else:
        c = np.maximum(cs,0); F = cs>0
        while True:
            cs, mu = eqp(F); it+=1
            if (cs[F] >= 0).all():
                c = cs
                viol = np.where(~F, -mu, -np.inf); i = viol.argmax()
                if viol[i] <= 1e-12*max(1.0, np.abs(p).max()): break
                F[i]=True
            else:
                dlt = cs - c
                ratio = np.where(F & (cs<0), c/np.maximum(-dlt,1e-300), np.inf)
                i = ratio.argmin(); alpha = min(1.0, ratio[i])
                c = c + alpha*dlt; c[i]=0; F[i]=False
            if it>=maxit: break
    f = G@c
    lam = np.linalg.solve(Ls.T, np.linalg.solve(Ls, rhs - Jb.T@f))
    a = -np.linalg.solve(La.T, w + Y@lam)
    return a, f, c, it
