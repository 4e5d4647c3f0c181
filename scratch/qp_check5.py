import sys; sys.path.insert(0,'/root/repo/scratch')
from qp_proto import *
from qp_proto2 import solve_c
import mpmath as mp
mp.mp.dps = 60
s = O.sim(); sl = O.sim(qp_long_double=1); x0,_,_ = s.get_state()
rng = np.random.default_rng(1)
def exact_eqp(d, F):
    # full 74-var KKT for the given free set, in 60-digit arithmetic
    H = mp.matrix(d['H'].tolist()); g = mp.matrix(d['g'].tolist())
    M6 = d['M'][:6,:]; C6 = d['C'][:6]; JT = d['JFeet'].T
    rows=[]; rhs=[]
    for r in range(6):
        row=[0.0]*74
        for c in range(30): row[c]=float(M6[r,c])
        for c in range(12): row[30+c]=-float(JT[r,c])
        rows.append(row); rhs.append(-float(C6[r]))
    for r in range(12):
        row=[0.0]*74; row[30+r]=-1.0
        for c in range(32): row[42+c]=float(G[r,c])
        rows.append(row); rhs.append(0.0)
    for i in range(32):
        if not F[i]:
            row=[0.0]*74; row[42+i]=1.0; rows.append(row); rhs.append(0.0)
    A = mp.matrix(rows); b = mp.matrix(rhs); m=A.rows
    K = mp.zeros(74+m, 74+m)
    K[0:74,0:74]=H; K[0:74,74:]= A.T; K[74:,0:74]=A
    sol = mp.lu_solve(K, mp.matrix(list(-g)+list(b)))
    x = np.array([float(sol[i]) for i in range(74)])
    lam = [sol[74+i] for i in range(m)]   # H x + g + A' lam = 0 ; for bound rows c_i=0 : multiplier mu_i = -lam? (c_i>=0 -> need lam<=0)
    mu = np.array([float(-l) for l in lam[18:]])
    return x, mu
wm=[0,0]; wo=[0,0]; wl=[0,0]
for trial in range(150):
    x = x0.copy(); scale = [0,0.01,0.05,0.1,0.3][trial%5]
    x[:30] += scale*rng.standard_normal(30)*0.3; x[30:] = scale*3*rng.standard_normal(30)
    s.set_state(x, 0.0, x[30:]); d = s.stage(x, 0.0)
    sl.set_state(x, 0.0, x[30:]); dl = sl.stage(x, 0.0)
    a,f,c,it = solve_c(d)
    F = c>0
    xe, mu = exact_eqp(d, F)
    ok = (xe[42:][F].min()>=0) and (len(mu)==0 or mu.min()>=-1e-9*abs(mu).max())
    rel = lambda u,v: np.abs(u-v).max()/max(1,np.abs(v).max())
    em=(rel(a,xe[:30]), rel(f,xe[30:42])); eo=(rel(d['x'][:30],xe[:30]), rel(d['x'][30:42],xe[30:42])); el=(rel(dl['x'][:30],xe[:30]), rel(dl['x'][30:42],xe[30:42]))
    wm=np.maximum(wm,em); wo=np.maximum(wo,eo); wl=np.maximum(wl,el)
    if not ok or max(em)>1e-8: print(trial, scale, 'certified' if ok else 'NOT OPT minc %.1e minmu %.1e / %.1e'%(xe[42:][F].min(), mu.min() if len(mu) else 0, abs(mu).max() if len(mu) else 0), 'mine a %.1e f %.1e | oracle %.1e %.1e | oracle-ld %.1e %.1e'%(em+eo+el))
print('worst mine', wm, 'oracle', wo, 'oracle-ld', wl)
