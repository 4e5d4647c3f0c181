import sys; sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import numpy as np, hostemu_py, orc_py
import linearmpchumanoid_b200 as lmh
wp = lmh.walk_plan(range(2))
O = orc_py.oracle()
for env in range(2):
    so = O.sim(time_horizon=1.6, sim_time=5.0)
    so.set_walk_plan(wp['zmp_x'][env], wp['zmp_y'][env], wp['phase_T'], wp['foot_coef'][env], wp['contact'][env])
    e = hostemu_py.Emu(horizon=1.6)
    e.set_walk_plan(wp['zmp_x'][env], wp['zmp_y'][env], wp['phase_T'], wp['foot_coef'][env], wp['contact'][env])
    x0,_,_ = so.get_state()
    n = 150
    lo = so.step(n)
    x,vh,t,lg,st = e.step(x0, np.zeros(30), 0.0, n)
    print('env',env,'L %.3f'%wp['step_len'][env],'x diff', np.abs(x-lo['x'][-1]).max(), 'com diff', np.abs(lg['com_k4']-lo['com_k4']).max(), 'rfoot diff', np.abs(lg['rfoot']-lo['rfoot']).max(),
          'k exact', np.array_equal(lg['k_index'],lo['k_index']), 'qp fail', st['qp_fail'], lo['qp_fail'].sum(), 'alive', st['alive'])
    print('  com x at 0.5,1.0,1.5 s', lo['com_k4'][[49,99,149],0], 'rfoot x', lo['rfoot'][[49,99,149],0], 'rfoot z max', lo['rfoot'][:,2].max(), 'lfoot x', lo['lfoot'][[49,99,149],0], 'base z', lo['x'][[49,99,149],2])
