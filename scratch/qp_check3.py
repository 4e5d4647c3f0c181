import sys; sys.path.insert(0,'/root/repo/scratch')
from qp_proto import *
s = O.sim(); x0,_,_ = s.get_state()
rng = np.random.default_rng(1)
worst = np.zeros(5)
for trial in range(300):
    x = x0.copy(); scale = [0,0.01,0.05,0.1,0.3][trial%5]
    x[:30] += scale*rng.standard_normal(30)*0.3; x[30:] = scale*3*rng.standard_normal(30)
    s.set_state(x, 0.0, x[30:]); d = s.stage(x, 0.0)
    a,f,c,it = solve(d)
    Ha = d['H'][:30,:30]; ga = d['g'][:30]; M6 = d['M'][:6,:]; C6 = d['C'][:6]; Jb = d['JFeet'][:,:6]
    # multipliers by least squares on the a-stationarity
    lam = np.linalg.lstsq(M6.T, -(Ha@a+ga), rcond=None)[0]
    st_a = np.abs(Ha@a+ga+M6.T@lam).max()/np.abs(ga).max()
    nu = wF*f - Jb@lam
    mu = eps*c + G.T@nu
    sc = max(1,np.abs(G.T@(Jb@lam)).max())
    r = np.array([st_a, np.abs(M6@a - Jb.T@f + C6).max()/max(1,np.abs(C6).max()), max(0,-mu.min())/sc, np.abs(mu*c).max()/sc/max(1,c.max()), max(0,-c.min())])
    worst = np.maximum(worst, r)
    if r.max()>1e-9: print(trial, scale, r, it)
print('worst [stat_a, eq, dual sign, compl, c<0]', worst)
