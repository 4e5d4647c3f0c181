import sys; sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scratch'); sys.path.insert(0,'/root/repo/oracle')
import numpy as np, hostemu_py, orc_py
from qp_proto3 import solve_rank, pchol6, Gf
from qp_proto import G
e=hostemu_py.Emu()
z=np.load('/tmp/fail_state.npz'); x,vh,t=z['x'],z['vh'],float(z['t'])
# replay the 4 stages via rk4 pieces to find the failing stage
dt=0.01
O=orc_py.oracle(); so=O.sim()
def stage(xx,vv,tt):
    d,vv2=e.stage(xx,vv,tt); return d,vv2
k=[]; xs=x; v=vh
for st,(cst,tt) in enumerate([(0,t),(0.5*dt,t+0.5*dt),(0.5*dt,t+0.5*dt),(dt,t+dt)]):
    xi = x if st==0 else x+cst*k[-1]
    d,v=stage(xi,v,tt); k.append(d['xdot'])
    print('stage',st,'iters',d['qp_iters'],'status',d['qp_status'],'nact',(d['c']==0).sum())
    if d['qp_status']:
        so.set_state(xi,tt,vh if st==0 else None); do=so.stage(xi,tt)
        a,f,c,it=solve_rank(do)
        print(' proto iters',it,'nact',(c==0).sum(), 'oracle iters', do['qp_iters'], 'f diff proto-oracle', np.abs(f-do['f']).max())
        np.savez('/tmp/fail_stage.npz',x=xi,vh=vh,t=tt)
