"""First GPU contact: parity of stage dumps + closed loop vs the oracle, and a timing sweep."""
import sys, time, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import numpy as np
import linearmpchumanoid_b200 as lmh
import orc_py
O = orc_py.oracle()
so = O.sim()
xo, _, _ = so.get_state()
print('fp64 peak', lmh.measure_fp64_peak(0))
for nt in (32, 64, 128):
    E = 64
    b = lmh.Batch(E, threads_per_env=nt, log_steps=100)
    q0, zc, m, it = b.nominal()
    print('nt', nt, 'q0 diff', np.abs(q0 - xo[:30]).max(), 'info', b.launch_info())
    rng = np.random.default_rng(0)
    X = np.tile(xo, (E, 1)); VH = np.zeros((E, 30)); T = np.zeros(E)
    for e in range(1, E):
        sc = [0.002, 0.01, 0.03][e % 3]
        X[e, :30] += sc * rng.standard_normal(30); X[e, 30:] = 3 * sc * rng.standard_normal(30)
        VH[e] = X[e, 30:] + 0.1 * sc * rng.standard_normal(30); T[e] = 0.01 * (e % 7)
    b.set_state(X, VH, T)
    d = b.stage()
    worst = {}
    for e in range(E):
        so.set_state(X[e], T[e], VH[e]); do = so.stage(X[e], T[e])
        for k, ko in (('M','M'),('C','C'),('Cg','Cg'),('AG','AG'),('JFeet','JFeet'),('Jpqp','Jpqp'),('CoM','CoM'),('xRef','xRef'),('Ha','H'),('ga','g'),('qpp','qpp'),('tau','tau'),('f','f'),('xdot','xdot')):
            a = do[ko]; 
            if ko == 'H': a = a[:30,:30]
            if ko == 'g': a = a[:30]
            err = np.abs(a - d[k][e]).max() / max(1.0, np.abs(a).max())
            worst[k] = max(worst.get(k, 0), err)
        assert d['k_index'][e] == do['k_index']
    print('  stage worst rel err', {k: float('%.1e' % v) for k, v in worst.items()}, 'iters', d['qp_iters'].max())
    # closed loop from the nominal posture, 100 steps, env 0 unpushed
    b.reset()
    b.step(100)
    lg = b.log(0, 100)
    so2 = O.sim(); lo = so2.step(100)
    x, vh, t = b.get_state()
    print('  closed loop: x', np.abs(x[0] - lo['x'][-1]).max(), 'com', np.abs(lg['com_k4'][:, 0] - lo['com_k4']).max(),
          'tau', np.abs(lg['tau_k1'][:, 0] - lo['tau_k1']).max(), 'k', np.abs(lg['k_index'][:, 0] - lo['k_index']).max(), 't', t[0])
    b.close()
    # timing at 4096 envs
    E = 4096
    b = lmh.Batch(E, threads_per_env=nt)
    ps, dv = lmh.push_plan(range(E))
    b.set_pushes(ps, dv)
    b.step(2)
    for n in (5, 20):
        ms = b.step_timed(n)
        print('  E=4096 nt=%d steps=%d: %.2f ms -> %.0f env-steps/s' % (nt, n, ms, E * n / ms * 1e3))
    st = b.stats()
    print('  stats: alive', st['alive'].sum(), 'qp_fail', st['qp_fail'].sum(), 'iters/stage', st['qp_iters'].sum() / (st['steps'].sum() * 4.0), 'max tau', st['max_abs_tau'].max())
    b.close()
