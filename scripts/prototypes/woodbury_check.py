import sys; sys.path.insert(0,'/root/repo/scratch'); sys.path.insert(0,'/root/repo/oracle')
import numpy as np, orc_py
from qp_common import G, eps, wF
q=np.load('/root/repo/tests/golden/qp_exact.npz')
O=orc_py.oracle(); s=O.sim()
rel=lambda a,b: np.abs(a-b).max()/max(1,np.abs(b).max())
worst=0
for i in range(len(q['g'])):
    x=q['state'][i]; s.set_state(x,0.0,x[30:]); d=s.stage(x,0.0); xe=q['x'][i]
    AG=d['AG']; J=d['JFeet']; M6=d['M'][:6]; C6=d['C'][:6]; ga=d['g'][:30]; Jb=J[:,:6]
    WJ=np.diag([10.]*6+[1.]*24)
    D=WJ+1e5*J.T@J; U=np.sqrt(4000.)*AG[3:6]          # Ha = D + U'U
    idx1=np.arange(18); idx2=np.arange(18,30)
    L1=np.linalg.cholesky(D[:18,:18])
    B=np.concatenate([M6, ga[None,:], U],0).T          # 30 x 10
    Yt=np.zeros((30,10)); Yt[:18]=np.linalg.solve(L1,B[:18]); Yt[18:]=B[18:]/1.0   # sqrt(wJ)=1
    Gm=Yt.T@Yt
    Cm=np.eye(3)+Gm[7:,7:]; Ci=np.linalg.inv(Cm)
    P=Gm[:7,:7]-Gm[:7,7:]@Ci@Gm[7:,:7]                  # [M6;ga] Ha^-1 [M6;ga]'
    Sa=P[:6,:6]; rhs=C6-P[:6,6]
    # QP in c with known active set from exact solution
    F=xe[42:]>1e-14
    Ls=np.linalg.cholesky(Sa); Wm=np.linalg.solve(Ls,Jb.T@G)
    Q=wF*G.T@G+Wm.T@Wm+eps*np.eye(32); p=Wm.T@np.linalg.solve(Ls,rhs)
    c=np.zeros(32); ii=np.where(F)[0]; c[ii]=np.linalg.solve(Q[np.ix_(ii,ii)],p[ii])
    f=G@c; lam=np.linalg.solve(Sa,rhs-Jb.T@f)
    # a = -Ha^-1 (ga + M6' lam)
    coef=np.concatenate([lam,[1.0]])                    # combination of columns [M6 rows, ga]
    z=Gm[7:,:7]@coef; cu=-Ci@z
    yt=Yt[:,:7]@coef+Yt[:,7:]@cu
    a=np.zeros(30); a[:18]=-np.linalg.solve(L1.T,yt[:18]); a[18:]=-yt[18:]/1.0
    e=(rel(a,xe[:30]),rel(f,xe[30:42])); worst=max(worst,*e)
    print(i,int(q['n_active'][i]),'a %.1e f %.1e'%e, 'cond C %.1e'%np.linalg.cond(Cm))
print('worst',worst)
