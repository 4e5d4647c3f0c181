import sys; sys.path.insert(0,'/root/repo/scratch')
from qp_proto import *
s = O.sim(); x0,_,_ = s.get_state()
rng = np.random.default_rng(1)
def obj(d,a,c):
    Ha = d['H'][:30,:30]; ga = d['g'][:30]; f=G@c
    return 0.5*a@Ha@a + ga@a + 0.5*wF*f@f + 0.5*eps*c@c
def feas(d,a,c):
    M6 = d['M'][:6,:]; C6 = d['C'][:6]; Jb = d['JFeet'][:,:6]
    return np.abs(M6@a - Jb.T@(G@c) + C6).max(), min(c.min(),0)
for trial in range(60):
    x = x0.copy(); scale = [0,0.01,0.05,0.1,0.3][trial%5]
    x[:30] += scale*rng.standard_normal(30)*0.3; x[30:] = scale*3*rng.standard_normal(30)
    s.set_state(x, 0.0, x[30:]); d = s.stage(x, 0.0)
    a,f,c,it = solve(d)
    xo = d['x']; ao, fo, co = xo[:30], xo[30:42], xo[42:]
    ef = np.abs(f-fo).max()/max(1,np.abs(fo).max())
    if ef>1e-8:
        print(trial, 'ef %.1e'%ef, 'obj mine %.12e oracle %.12e diff %.3e'%(obj(d,a,c), obj(d,ao,co), obj(d,a,c)-obj(d,ao,co)), 'feas mine', feas(d,a,c), 'oracle', feas(d,ao,co), 'f-Gc oracle %.1e'%np.abs(fo-G@co).max())
