import sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo/oracle')
import orc_py
f64=C.c_double; i32=C.c_int32
_D=[("M",900),("C",30),("Cg",30),("AG",180),("AGpqp",6),("Jpqp",12),("JFeet",360),("CoM",3),("comVel",3),("comAngMom",3),("xRef",3),("yRef",3),
    ("qppRef",30),("hGpRef",6),("footAccRef",12),("Ha",900),("ga",30),("a",30),("f",12),("c",32),("lam",6),("qpp",30),("tau",24),("T7",12),("T14",12),("xdot",60)]
class Dump(C.Structure):
    _fields_=[(n,f64*k) for n,k in _D]+[("k_index",i32),("qp_status",i32),("qp_iters",i32),("pad",i32)]
    def d(self):
        o={n:np.ctypeslib.as_array(getattr(self,n)).copy() for n,_ in _D}
        for n,s in (("M",(30,30)),("AG",(6,30)),("JFeet",(12,30)),("Ha",(30,30))): o[n]=o[n].reshape(s)
        o.update(k_index=self.k_index,qp_status=self.qp_status,qp_iters=self.qp_iters); return o
class StepLog(C.Structure):
    _fields_=[("com_k4",f64*3),("rfoot",f64*3),("lfoot",f64*3),("tau_k1",f64*24),("k_index",i32*4),("qp_iters",i32),("qp_fail",i32)]
class Stats(C.Structure):
    _fields_=[("max_abs_tau",f64),("max_com_err",f64),("max_abs_qpp",f64),("qp_iters",i32),("qp_fail",i32),("steps",i32),("alive",i32)]
L=C.CDLL('/root/repo/tests/hostemu/libhostemu.so')
L.emu_create.restype=C.c_void_p; L.emu_create.argtypes=[f64,f64,f64,C.c_void_p]
L.emu_nominal.argtypes=[C.c_void_p]*5
L.emu_stage.argtypes=[C.c_void_p,C.c_void_p,C.c_void_p,f64,C.POINTER(Dump)]
L.emu_step.argtypes=[C.c_void_p,C.c_void_p,C.c_void_p,C.c_void_p,i32,C.c_void_p,C.c_void_p]
def P(a): return a.ctypes.data_as(C.c_void_p)
class Emu:
    def __init__(s,dt=0.01,T=5.0,hz=0.5): s.h=L.emu_create(dt,T,hz,None)
    def nominal(s):
        q=np.zeros(30); z=np.zeros(1); m=np.zeros(1); it=np.zeros(1,np.int32); L.emu_nominal(s.h,P(q),P(z),P(m),P(it)); return q,z[0],m[0],int(it[0])
    def stage(s,x,vh,t):
        d=Dump(); vh=vh.copy(); L.emu_stage(s.h,P(np.ascontiguousarray(x)),P(vh),t,C.byref(d)); return d.d(),vh
    def step(s,x,vh,t,n):
        x=x.copy(); vh=vh.copy(); tt=np.array([t]); logs=(StepLog*n)(); st=Stats(); L.emu_step(s.h,P(x),P(vh),P(tt),n,logs,C.byref(st)); return x,vh,tt[0],logs,st
def compare(do,de,keys=None,verbose=True):
    worst=0
    m={'H':'Ha','g':'ga'}
    for k in ['M','C','Cg','AG','AGpqp','Jpqp','JFeet','CoM','comVel','comAngMom','xRef','yRef','qppRef','hGpRef','footAccRef','H','g','qpp','tau','f','xdot']:
        a=do[k]; b=de[m.get(k,k)]
        if k=='H': a=a[:30,:30]
        if k=='g': a=a[:30]
        err=np.abs(a-b).max()/max(1e-12,np.abs(a).max()) if np.abs(a).max()>1e-9 else np.abs(a-b).max()
        worst=max(worst,err)
        if verbose: print(f'{k}:{err:.1e}',end=' ')
    for k,(a,b) in {'T7R':(do['T7'][:3,:3].ravel(),de['T7'][:9]),'T7p':(do['T7'][:3,3],de['T7'][9:]),'a':(do['x'][:30],de['a']),'c':(do['x'][42:],de['c'])}.items():
        err=np.abs(a-b).max()/max(1e-12,np.abs(a).max()); 
        if verbose: print(f'{k}:{err:.1e}',end=' ')
    if verbose: print('| k',do['k_index'],de['k_index'],'iters',do['qp_iters'],de['qp_iters'],de['qp_status'])
    return worst
if __name__=='__main__':
    O=orc_py.oracle(); so=O.sim(); e=Emu()
    q0,z,m,it=e.nominal(); xo,_,_=so.get_state()
    print('q0 diff',np.abs(q0-xo[:30]).max(),'zcom',z-so.zcom,'mass',m-so.mass,'ik',it,so.ik_iterations)
    rng=np.random.default_rng(0)
    for trial in range(6):
        x=xo.copy(); vh=np.zeros(30)
        if trial:
            sc=[0,0.002,0.01,0.05,0.1,0.3][trial]
            x[:30]+=sc*rng.standard_normal(30); x[30:]=3*sc*rng.standard_normal(30); vh=x[30:]+0.1*sc*rng.standard_normal(30)
        so.set_state(x,0.013*trial,vh); do=so.stage(x,0.013*trial)
        de,vh2=e.stage(x,vh,0.013*trial)
        compare(do,de)
