"""T(n): time of one wave with n one-warp environments per SM (E = 148 n, LMH_EPB = n, one block per SM)."""
import os, subprocess, sys
sys.path.insert(0, '/root/repo')
if len(sys.argv) > 1:
    n = int(sys.argv[1])
    import linearmpchumanoid_b200 as lmh
    E = 148 * n
    b = lmh.Batch(E)
    step, dv = lmh.push_plan(range(E))
    b.set_pushes(step, dv)
    b.step(100)
    ms = b.step_timed(100)
    print('n %2d  E %5d  %.1f us/step  %.0f env-steps/s' % (n, E, ms * 10, E * 100 / ms * 1e3), b.launch_info()['blocks'])
else:
    for n in range(1, 11):
        subprocess.run([sys.executable, __file__, str(n)], env=dict(os.environ, LMH_EPB=str(n)))
