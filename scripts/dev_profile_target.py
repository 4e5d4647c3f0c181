import sys
sys.path.insert(0, '/root/repo')
import linearmpchumanoid_b200 as lmh
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
b = lmh.Batch(E)
step, dv = lmh.push_plan(range(E))
b.set_pushes(step, dv)
b.step(100)           # launch 0
ms = b.step_timed(10) # launch 1: profiled -- settled regime
st = b.stats()
print('E', E, 'nt', b.launch_info()['threads_per_env'], 'ms', ms, 'env-steps/s', E * 10 / ms * 1e3, 'fallback stages', st['qp_fallback'].sum(), 'iters/stage', st['qp_iters'].sum()/(4.0*st['steps'].sum()))
