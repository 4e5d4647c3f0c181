import sys; sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import numpy as np, hostemu_py
import linearmpchumanoid_b200 as lmh
e=hostemu_py.Emu(); q0,_,_,_=e.nominal()
lo,hi=int(sys.argv[1]),int(sys.argv[2])
step,dv=lmh.push_plan(range(hi))
for env in range(lo,hi):
    x=np.concatenate([q0,np.zeros(30)]); vh=np.zeros(30); t=0.0
    for s in range(70):
        if s==step[env]: x[30:36]+=dv[env]
        x0,vh0,t0=x.copy(),vh.copy(),t
        x,vh,t,lg,st=e.step(x,vh,t,1)
        if st['qp_fail']:
            print('FAIL env',env,'step',s); np.savez('/tmp/fail_state.npz',x=x0,vh=vh0,t=t0); sys.exit(0)
print('none in',lo,hi)
