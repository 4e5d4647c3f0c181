import sys
sys.path.insert(0, '/root/repo')
import linearmpchumanoid_b200 as lmh
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
b = lmh.Batch(E)
step, dv = lmh.push_plan(range(E))
b.set_pushes(step, dv)
b.step(15)            # launch 0
ms = b.step_timed(20) # launch 1: profiled -- pushes arriving (steps 15..35)
st = b.stats()
print('E', E, 'ms', ms, 'env-steps/s', E * 20 / ms * 1e3, 'iters/stage', st['qp_iters'].sum()/(4.0*st['steps'].sum()))
