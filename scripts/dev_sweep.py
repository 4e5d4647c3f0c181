import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import linearmpchumanoid_b200 as lmh
E = 4096
for nt in [int(a) for a in sys.argv[1:]] or [32, 64]:
    b = lmh.Batch(E, threads_per_env=nt)
    step, dv = lmh.push_plan(range(E))
    b.set_pushes(step, dv)
    b.step(2)
    ms20 = b.step_timed(58)     # steps 2..60: pushes happen here
    ms = b.step_timed(140)      # settled
    st = b.stats()
    print('nt %d: push phase %.0f env-steps/s, settled %.0f env-steps/s, overall %.0f ; iters/stage %.3f fails %d' % (
        nt, E * 58 / ms20 * 1e3, E * 140 / ms * 1e3, E * 198 / (ms + ms20) * 1e3, st['qp_iters'].sum() / (st['steps'].sum() * 4.0), st['qp_fail'].sum()), b.launch_info())
    b.close()
