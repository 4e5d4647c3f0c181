"""Prototype of the reduced WBC QP solve (y-formulation primal active set) vs the oracle's GI solver."""
import sys; sys.path.insert(0,'/root/repo/oracle')
import orc_py, numpy as np
np.set_printoptions(linewidth=200, precision=6)
O = orc_py.oracle()
mu=0.7
Bf = np.array([[mu,0,1],[0,mu,1],[-mu,0,1],[0,-mu,1]]).T  # 3x4
verts = np.array([[0.1,0.025,0],[0.1,-0.025,0],[-0.05,0.025,0],[-0.05,-0.025,0]])
def cross(p): return np.array([[0,-p[2],p[1]],[p[2],0,-p[0]],[-p[1],p[0],0]])
Gfoot = np.zeros((6,16))
for v in range(4):
    Gfoot[0:3,4*v:4*v+4] = cross(verts[v])@Bf
    Gfoot[3:6,4*v:4*v+4] = Bf
G = np.zeros((12,32)); G[0:6,0:16]=Gfoot; G[6:12,16:32]=Gfoot
eps=1e-8; wF=1.0

def solve(d, verbose=False):
    Ha = d['H'][:30,:30]; ga = d['g'][:30]
    M6 = d['M'][:6,:]; C6 = d['C'][:6]; Jb = d['JFeet'][:,:6]
    La = np.linalg.cholesky(Ha)
    Y = np.linalg.solve(La, M6.T); w = np.linalg.solve(La, ga)
    Sa = Y.T@Y; rhs = C6 - Y.T@w
    Z = wF*Sa + Jb.T@Jb; Lz = np.linalg.cholesky(Z)
    V = np.linalg.solve(Lz, Jb.T)
    Kt = np.eye(12) - V.T@V
    b = V.T@np.linalg.solve(Lz, rhs)
    def eqp(F):
        GF = G[:,F]; A = GF@GF.T + (eps/wF)*Kt
        y = np.linalg.solve(A, b)   # (cholesky in the kernel)
        return y, G.T@y
    F = np.ones(32,bool); y, gy = eqp(F); it=1
    if (gy>=0).all():
        c = gy
    else:
        c = np.maximum(gy,0); F = gy>0
        while True:
            y, gy = eqp(F); it+=1
            cs = np.where(F, gy, 0.0)
            if (cs[F] >= -1e-12*max(1,abs(cs).max())).all():
                c = np.maximum(cs,0)
                viol = np.where(~F, gy, -np.inf)
                i = viol.argmax()
                if viol[i] <= 1e-10*max(1.0,abs(cs).max()): break
                F[i]=True
            else:
                dlt = cs - c
                ratio = np.where(F & (dlt<0), c/np.maximum(-dlt,1e-300), np.inf)
                i = ratio.argmin(); alpha = min(1.0, ratio[i])
                c = c + alpha*dlt; c[i]=0; F[i]=False
            if it>100: print('ITER CAP'); break
    f = G@c
    lam = np.linalg.solve(Sa, rhs - Jb.T@f)
    a = -np.linalg.solve(La.T, w + Y@lam)
    return a, f, c, it

if __name__ == '__main__':
    s = O.sim(); x0,_,_ = s.get_state()
    rng = np.random.default_rng(1)
    worst=0
    for trial in range(300):
        x = x0.copy(); scale = [0,0.01,0.05,0.1,0.3][trial%5]
        x[:30] += scale*rng.standard_normal(30)*0.3; x[30:] = scale*3*rng.standard_normal(30)
        s.set_state(x, 0.0, x[30:]); d = s.stage(x, 0.0)
        a,f,c,it = solve(d)
        xo = d['x']; ao, fo, co = xo[:30], xo[30:42], xo[42:]
        ea = np.abs(a-ao).max()/max(1,np.abs(ao).max()); ef = np.abs(f-fo).max()/max(1,np.abs(fo).max())
        nact = (co<1e-9).sum()
        worst=max(worst,ea,ef)
        if ea>1e-8 or ef>1e-8:
            print(trial, 'scale',scale,'oracle iters', d['qp_iters'], 'st', d['qp_status'], 'mine', it, 'nact', nact, (c<1e-12).sum(), 'err a %.2e f %.2e'%(ea,ef), 'kkt %.1e'%d['kkt'][0], '|a|%.1e'%np.abs(ao).max())
    print('worst', worst)
