import sys; sys.path.insert(0,'/root/repo/scratch')
from qp_proto import *
s = O.sim(); x0,_,_ = s.get_state()
sl = O.sim(qp_long_double=1)
rng = np.random.default_rng(1)
def feas(d,a,c):
    M6 = d['M'][:6,:]; C6 = d['C'][:6]; Jb = d['JFeet'][:,:6]
    return np.abs(M6@a - Jb.T@(G@c) + C6).max(), min(c.min(),0)
w=[0,0,0]
for trial in range(100):
    x = x0.copy(); scale = [0,0.01,0.05,0.1,0.3][trial%5]
    x[:30] += scale*rng.standard_normal(30)*0.3; x[30:] = scale*3*rng.standard_normal(30)
    s.set_state(x, 0.0, x[30:]); d = s.stage(x, 0.0)
    sl.set_state(x, 0.0, x[30:]); dl = sl.stage(x, 0.0)
    a,f,c,it = solve(d)
    xo = d['x']; xl = dl['x']
    e_d = np.abs(f-xo[30:42]).max()/max(1,np.abs(f).max()); e_l = np.abs(f-xl[30:42]).max()/max(1,np.abs(f).max())
    ea_d = np.abs(a-xo[:30]).max()/max(1,np.abs(a).max()); ea_l = np.abs(a-xl[:30]).max()/max(1,np.abs(a).max())
    if e_l>1e-9 or ea_l>1e-9: print(trial, scale, 'f: vs double %.1e vs longdouble %.1e ; a: %.1e %.1e'%(e_d,e_l,ea_d,ea_l), 'ld feas', feas(dl, xl[:30], xl[42:]), dl['qp_status'], dl['qp_iters'], it)
    w=[max(w[0],e_l),max(w[1],ea_l)]
print(w)
