import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import linearmpchumanoid_b200 as lmh
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
b = lmh.Batch(E, time_horizon=1.6)
wp = lmh.walk_plan(range(E))
b.set_walk_plans(wp["zmp_x"], wp["zmp_y"], wp["phase_T"], wp["foot_coef"], wp["contact"], plan_ids=np.arange(E))
b.step(60)            # launch 0
ms = b.step_timed(20) # launch 1: profiled (steps 60..80: a swing phase)
st = b.stats()
print('E', E, 'ms', ms, 'env-steps/s', E * 20 / ms * 1e3, 'iters/stage', st['qp_iters'].sum()/(4.0*st['steps'].sum()), 'alive', st['alive'].mean())
