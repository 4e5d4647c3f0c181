"""rank-revealing wrench-space primal active set (prototype of the kernel algorithm)"""
import sys; sys.path.insert(0,'/root/repo/scratch'); sys.path.insert(0,'/root/repo/oracle')
import numpy as np, orc_py
from qp_common import G, eps, wF
Gf=G[:6,:16]
def pchol6(P, tol=1e-10):
    """diagonally pivoted Cholesky of a 6x6 PSD matrix -> B (6x6, zero columns beyond the rank) with P = B B'"""
    S=P.copy(); B=np.zeros((6,6)); done=np.zeros(6,bool); dmax=max(np.diag(P).max(),1e-300)
    for k in range(6):
        dd=np.where(done,-1,np.diag(S)); p=dd.argmax()
        if dd[p] <= tol*dmax: break
        col=S[:,p]/np.sqrt(dd[p]); col[done]=0.0   # rows already pivoted have zero residual
        B[:,k]=col; S=S-np.outer(col,col); done[p]=True
    return B
def solve_rank(d, maxit=200):
    Ha=d['H'][:30,:30]; ga=d['g'][:30]; M6=d['M'][:6]; C6=d['C'][:6]; Jb=d['JFeet'][:,:6]
    La=np.linalg.cholesky(Ha); Y=np.linalg.solve(La,M6.T); w=np.linalg.solve(La,ga)
    Sa=Y.T@Y; rhs=C6-Y.T@w; Ls=np.linalg.cholesky(Sa); sv=np.linalg.solve(Ls,rhs)
    def eqp(F):
        B=np.zeros((12,12))
        for f in range(2):
            m=F[16*f:16*f+16]; P=(Gf[:,m]@Gf[:,m].T) if m.any() else np.zeros((6,6))
            B[6*f:6*f+6,6*f:6*f+6]=pchol6(P)
        T=B.T@B; W=np.linalg.solve(Ls,Jb.T@B)
        A=wF*T+W.T@W+eps*np.eye(12)
        zeta=np.linalg.solve(A,W.T@sv)
        f=B@zeta
        Tp=T+np.diag((np.diag(T)==0).astype(float))
        yv=B@np.linalg.solve(Tp,zeta)
        cs=np.where(F,G.T@yv,0.0)
        lam=np.linalg.solve(Sa,rhs-Jb.T@f)
        mu=G.T@(wF*f-Jb@lam)        # true multipliers on W (>= 0 at optimum)
        return cs,mu,f,lam
    F=np.ones(32,bool); cs,mu,f,lam=eqp(F); it=1
    if (cs>=0).all(): c=cs
    else:
        c=np.maximum(cs,0); F=cs>0
        while True:
            cs,mu,f,lam=eqp(F); it+=1
            if (cs[F]>=0).all():
                c=cs; sc=max(1.0,np.abs(G.T@(Jb@lam)).max())
                viol=np.where(~F,mu,np.inf); i=viol.argmin()
                if viol[i]>=-1e-10*sc: break
                F[i]=True
            else:
                dl=cs-c; r=np.where(F&(cs<0),c/np.maximum(-dl,1e-300),np.inf); i=r.argmin()
                c=c+min(1.0,r[i])*dl; c[i]=0; F[i]=False
            if it>=maxit: print('ITER CAP'); break
    f=G@c; lam=np.linalg.solve(Sa,rhs-Jb.T@f); a=-np.linalg.solve(La.T,w+Y@lam)
    return a,f,c,it
if __name__=='__main__':
    q=np.load('/root/repo/tests/golden/qp_exact.npz'); O=orc_py.oracle(); s=O.sim()
    rel=lambda a,b: np.abs(a-b).max()/max(1,np.abs(b).max()); worst=0
    for i in range(len(q['g'])):
        x=q['state'][i]; s.set_state(x,0.0,x[30:]); d=s.stage(x,0.0); xe=q['x'][i]
        a,f,c,it=solve_rank(d); e=(rel(a,xe[:30]),rel(f,xe[30:42])); worst=max(worst,*e)
        print(i,int(q['n_active'][i]),'iters',it,'a %.1e f %.1e'%e)
    print('worst',worst)
