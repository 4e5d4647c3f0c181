#!/usr/bin/env python3
"""Developer tool (run on the GPU box: PYTHONPATH=. python scripts/cmp_kernels.py): the opt-in latency kernel
(envs_per_block = -1) against the throughput kernel on a pushed batch -- determinism of each, the first step and
environment at which they differ (rounding: they are two kernels), iteration counts."""
import numpy as np, linearmpchumanoid_b200 as lmh
E,S=40,50
step,dv=lmh.push_plan(range(E),max_speed=0.5,max_step=20)
def run(kw, S):
    b=lmh.Batch(E,log_steps=S,**kw); b.set_pushes(step,dv); b.step(S)
    x=b.get_state()[0]; lg=b.log(0,S); st=b.stats(); b.close(); return x,lg,st
xa,la,sa=run({"envs_per_block":-1},S); xb,lb,sb=run({"envs_per_block":-1},S); xc,lc,sc=run({},S)
print("latency kernel deterministic:", np.array_equal(xa,xb))
d=np.abs(xa-xc).max(axis=1); bad=np.where(d>0)[0]
print("envs differing:", bad, "their push steps", step[bad], "qp_iters helpers/multi", sa["qp_iters"][bad], sc["qp_iters"][bad])
for e in bad[:3]:
    dt=np.abs(la["tau_k1"][:,e]-lc["tau_k1"][:,e]).max(axis=1); first=np.where(dt>0)[0]
    print(" env",e,"push at",step[e],"first differing step",first[:3], "qp_iters per step around", la["qp_iters"][max(0,first[0]-2):first[0]+3,e], lc["qp_iters"][max(0,first[0]-2):first[0]+3,e], "dtau", dt[first[0]])
