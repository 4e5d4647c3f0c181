import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import linearmpchumanoid_b200 as lmh
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
S = int(sys.argv[2]) if len(sys.argv) > 2 else 20
nt = int(sys.argv[3]) if len(sys.argv) > 3 else 0
b = lmh.Batch(E, threads_per_env=nt)
step, dv = lmh.push_plan(range(E), max_step=10)
b.set_pushes(step, dv)
b.step(2)
ms = b.step_timed(S)
print('E', E, 'S', S, 'nt', b.launch_info()['threads_per_env'], 'ms', ms, 'env-steps/s', E * S / ms * 1e3)
