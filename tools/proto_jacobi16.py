"""NumPy model of the rotation schedule of csrc/pd_stage_a_sym16.cuh: two-sided Jacobi on a 16 x 16 symmetric matrix
spread over 8 lanes (two columns each), odd-even transposition ordering, row indices relative to the lane's first
column; after every step the frame moves on by one column (the lane keeps its second column and takes the first one
of its right neighbour), the pair that closes the ring in odd steps is rotated by a quarter turn (a sign flip).
Axis 0 of every array is the lane.
    python tools/proto_jacobi16.py      # random SPD matrices: sweeps, residual, orthogonality, eigenvalue error"""
import numpy as np
N = 16
NL = 8


def rot_params(app, aqq, apq, frozen):
    tiny = frozen | (apq * apq <= 1e-36 * np.abs(app * aqq))
    d = aqq - app
    a2 = 2 * apq
    h = d * d + a2 * a2
    den = np.abs(d) + np.sqrt(h)
    tn = np.where(tiny, 0.0, np.where(d >= 0, a2, -a2) / np.where(den > 0, den, 1.0))
    c = 1 / np.sqrt(tn * tn + 1)
    return c, tn * c, tn, tiny


def step(A,B,Wa,Wb,idle,frozen):
    lam=np.arange(NL)
    app=A[:,0].copy(); aqq=B[:,1].copy(); apq=B[:,0].copy()
    c,s,tn,tiny=rot_params(app,aqq,apq,frozen)
    c=np.where(idle,0.0,c); s=np.where(idle,1.0,s)
    def colrot(X,Y):
        return s[:,None]*X + c[:,None]*Y, c[:,None]*X - s[:,None]*Y
    A2,B2=colrot(A,B); Wa2,Wb2=colrot(Wa,Wb)
    for X in (A2,B2):
        for j in range(1,8):
            src=(lam+j)&7; cj=c[src]; sj=s[src]
            rp=X[:,2*j].copy(); rq=X[:,2*j+1].copy()
            X[:,2*j]=sj*rp+cj*rq; X[:,2*j+1]=cj*rp-sj*rq
    tpq=np.where(tiny,apq,0.0)
    A2[:,0]=np.where(idle,app,aqq+tn*apq); A2[:,1]=np.where(idle,-apq,tpq)
    B2[:,0]=np.where(idle,-apq,tpq); B2[:,1]=np.where(idle,aqq,app-tn*apq)
    return A2,B2,Wa2,Wb2
def shift(A,B,Wa,Wb):
    lam=np.arange(NL); src=(lam+1)&7
    nA=np.empty_like(A); nB=np.empty_like(B)
    for r in range(16):
        nA[:,r]=B[:,(r+1)&15]; nB[:,r]=A[src,(r-1)&15]
    return nA,nB,Wb.copy(),Wa[src].copy()
def jacobi(T,maxsweeps=14):
    lam=np.arange(NL)
    A=np.stack([np.roll(T[:,2*l],-2*l) for l in lam]); B=np.stack([np.roll(T[:,2*l+1],-2*l) for l in lam])
    I=np.eye(N); Wa=np.stack([I[:,2*l] for l in lam]); Wb=np.stack([I[:,2*l+1] for l in lam])
    for sw in range(maxsweeps):
        off=(np.abs(A[:,1:]).sum()+np.abs(B[:,0]).sum()+np.abs(B[:,2:]).sum()); diag=np.abs(A[:,0]).sum()+np.abs(B[:,1]).sum()
        if off<=2e-17*diag: break
        for st in range(16):
            idle=((st&1)==1)&(lam==((7-(st>>1))&7))
            A,B,Wa,Wb=step(A,B,Wa,Wb,idle,False)
            A,B,Wa,Wb=shift(A,B,Wa,Wb)
    lamv=np.concatenate([[A[l,0],B[l,1]] for l in lam]); W=np.stack([x for l in lam for x in (Wa[l],Wb[l])],axis=1)
    return lamv,W,sw
if __name__ == "__main__":
  rng=np.random.default_rng(1)
  for trial in range(5):
    M=rng.standard_normal((N,N)); T=M@M.T+np.diag(rng.uniform(0,100,N))
    lamv,W,sw=jacobi(T)
    print(sw, np.abs(W.T@T@W-np.diag(lamv)).max()/np.abs(T).max(), np.abs(W.T@W-np.eye(N)).max(), np.abs(np.sort(lamv)-np.linalg.eigvalsh(T)).max())
