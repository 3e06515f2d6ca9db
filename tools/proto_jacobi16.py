"""NumPy model of the rotation schedule of csrc/pd_stage_a_sym16.cuh: two-sided Jacobi on a 16 x 16 symmetric matrix
spread over 8 lanes (two columns each), odd-even transposition ordering, row indices relative to the lane's first
column, frame changes between even and odd steps as shifts + renamings.  Axis 0 of every array is the lane.
    python tools/proto_jacobi16.py      # random SPD matrices: sweeps, residual, orthogonality, eigenvalue error"""
import numpy as np
N=16; NL=8
def rot_params(app,aqq,apq,frozen):
    tiny = frozen | (apq*apq <= 1e-36*np.abs(app*aqq))
    d=aqq-app; a2=2*apq; h=d*d+a2*a2
    den=np.abs(d)+np.sqrt(h)
    tn=np.where(tiny,0.0,np.where(d>=0,a2,-a2)/np.where(den>0,den,1.0))
    c=1/np.sqrt(tn*tn+1); s=tn*c
    return c,s,tn,tiny
def step(A,B,Wa,Wb,odd,frozen):
    """A,B: [NL][16] columns (slot0, slot1) in the lane's current frame (relative rows); local pair at rho=0,1."""
    lam=np.arange(NL)
    idle = odd & (lam==7)
    app=A[:,0]; aqq=B[:,1]; apq=B[:,0]   # T_pq = element row p of column q
    c,s,tn,tiny=rot_params(app,aqq,apq,frozen)
    c=np.where(idle,1.0,c); s=np.where(idle,0.0,s); tn=np.where(idle,0.0,tn); tiny=tiny|idle
    # column rotation (+ swap unless idle)
    def colrot(X,Y):
        newp=c[:,None]*X - s[:,None]*Y
        newq=s[:,None]*X + c[:,None]*Y
        X2=np.where(idle[:,None],newp,newq); Y2=np.where(idle[:,None],newq,newp)
        return X2,Y2
    A2,B2=colrot(A,B); Wa2,Wb2=colrot(Wa,Wb)
    # row rotations for j=1..7 on rows (2j,2j+1) with (c,s) of lane (lam+j)&7
    for X in (A2,B2):
        for j in range(1,8):
            src=(lam+j)&7
            cj=c[src]; sj=s[src]; noswap=idle[src]
            rp=X[:,2*j].copy(); rq=X[:,2*j+1].copy()
            newp=cj*rp-sj*rq; newq=sj*rp+cj*rq
            X[:,2*j]=np.where(noswap,newp,newq); X[:,2*j+1]=np.where(noswap,newq,newp)
    # own block, exact formulas.  after swap: slot0 = new q, slot1 = new p
    tpp=app-tn*apq; tqq=aqq+tn*apq; tpq=np.where(tiny,apq,0.0)
    # not idle: A2 is column new_q: rows (rho0 = pos of slot0 = new q row, rho1 = new p row)
    A2[:,0]=np.where(idle,app,tqq); A2[:,1]=np.where(idle,A[:,1],tpq)
    B2[:,0]=np.where(idle,B[:,0],tpq); B2[:,1]=np.where(idle,aqq,tpp)
    return A2,B2,Wa2,Wb2
def to_odd(A,B,Wa,Wb):
    lam=np.arange(NL); src=(lam+1)&7
    nA=np.empty_like(A); nB=np.empty_like(B)
    for r in range(16):
        nA[:,r]=B[:,(r+1)&15]           # own B, origin shifts by +1
        nB[:,r]=A[src,(r-1)&15]         # A of lane+1, origin 2lam+2 -> 2lam+1
    return nA,nB,Wb.copy(),Wa[src].copy()
def to_even(S0,S1,W0,W1):
    lam=np.arange(NL); src=(lam-1)&7
    nA=np.empty_like(S0); nB=np.empty_like(S1)
    for r in range(16):
        nA[:,r]=S1[src,(r+1)&15]
        nB[:,r]=S0[:,(r-1)&15]
    return nA,nB,W1[src].copy(),W0.copy()
def jacobi(T,maxsweeps=12):
    lam=np.arange(NL)
    A=np.stack([np.roll(T[:,2*l],-2*l) for l in lam]); B=np.stack([np.roll(T[:,2*l+1],-2*l) for l in lam])
    I=np.eye(N); Wa=np.stack([I[:,2*l] for l in lam]); Wb=np.stack([I[:,2*l+1] for l in lam])
    frozen=False
    for sw in range(maxsweeps):
        off=(np.abs(A[:,1:]).sum()+np.abs(B[:,0]).sum()+np.abs(B[:,2:]).sum())/2
        diag=np.abs(A[:,0]).sum()+np.abs(B[:,1]).sum()
        if off<=1e-17*diag: break
        for st in range(8):
            A,B,Wa,Wb=step(A,B,Wa,Wb,False,frozen)
            A,B,Wa,Wb=to_odd(A,B,Wa,Wb)
            A,B,Wa,Wb=step(A,B,Wa,Wb,True,frozen)
            A,B,Wa,Wb=to_even(A,B,Wa,Wb)
    lamv=np.concatenate([[A[l,0],B[l,1]] for l in lam]); W=np.stack([x for l in lam for x in (Wa[l],Wb[l])],axis=1)
    return lamv,W,sw
if __name__=="__main__":
    rng=np.random.default_rng(1)
    for trial in range(5):
        M=rng.standard_normal((N,N)); T=M@M.T+np.diag(rng.uniform(0,100,N))
        lamv,W,sw=jacobi(T)
        print(sw, np.abs(W.T@T@W-np.diag(lamv)).max()/np.abs(T).max(), np.abs(W.T@W-np.eye(N)).max(), np.abs(np.sort(lamv)-np.linalg.eigvalsh(T)).max())
