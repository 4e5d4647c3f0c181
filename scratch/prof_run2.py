import sys
sys.path.insert(0, '/root/repo')
import linearmpchumanoid_b200 as lmh
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 0
b = lmh.Batch(E, threads_per_env=nt)
step, dv = lmh.push_plan(range(E))
b.set_pushes(step, dv)
b.step(40)            # launch 0: most pushes have happened
ms = b.step_timed(10) # launch 1: profiled -- mixed regime (some envs still recovering)
print('E', E, 'nt', b.launch_info()['threads_per_env'], 'ms', ms, 'env-steps/s', E * 10 / ms * 1e3)
