import sys; sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scratch')
import numpy as np, hostemu_py, collections
import linearmpchumanoid_b200 as lmh
from qp_proto import G
e=hostemu_py.Emu()
q0,_,_,_=e.nominal()
N=60
step,dv=lmh.push_plan(range(N))
pat=collections.Counter(); tot=0; nfb=0
for env in range(N):
    x=np.concatenate([q0,np.zeros(30)]); vh=np.zeros(30); t=0.0
    for s in range(70):
        if s==step[env]: x[30:36]+=dv[env]
        # emulate rk4 by stages to see the c patterns: use step() for state, stage() for pattern at k1
        d,_=e.stage(x,vh,t)
        c=d['c']
        fr=[(c[16*f:16*f+16]>0) for f in range(2)]
        rk=[np.linalg.matrix_rank(G[6*f:6*f+6,16*f:16*f+16][:,fr[f]]) if fr[f].any() else 0 for f in range(2)]
        nz=[int(fr[f].sum()) for f in range(2)]
        if d['qp_iters']>1: pat[(tuple(rk),tuple(nz))]+=1
        tot+=1
        x,vh,t,_,st=e.step(x,vh,t,1)
        nfb+=st['qp_fallback']
print('stages',tot,'fallback stages (all 4 stages/step)',nfb)
for k,n in pat.most_common(20): print(n,k)
