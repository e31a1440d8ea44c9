"""NumPy emulation of the blocked factorisation of gp_kernel.cuh with ADAPTIVE refinement of the 8 x 8 tile solves: a solve
through the explicit pivot-tile inverse is refined only where the estimate max|L_pp| max|W_pp| (each rounded up to a power of
two, what the kernel reads from the exponent fields) exceeds tau.  Random light curves as in tools/parity_campaign.py, every
evaluation against the 80-bit value: always refine (tau = 0) / tau = 4, 8, 16 / never.  The kernel uses tau = 8.
usage: python tools/emulate_adaptive.py [seed] [cases]      Test infrastructure (imports oracle/); not product code."""
import os, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import gp_oracle as o
from gprotation_b200.synth import synthetic_lightcurve
import emulate_trsm as E
from scipy.linalg import solve_triangular

def est(Lpp,Wpp):
    ml=np.abs(Lpp).max(); mw=np.abs(Wpp).max()
    return 2.0**((np.floor(np.log2(ml))+1)+(np.floor(np.log2(mw))+1))

def lnlike_gpu(K,y,bs,tau):
    """tau: refine tile solves only where the estimate exceeds tau (tau=0: always, tau=inf: never)"""
    n=K.shape[0]; T=(n+bs-1)//bs
    L=np.zeros_like(K); yt=y.copy(); quad=logdet=0.0; nref=ntot=0
    for J in range(T):
        j0,j1=J*bs,min(n,(J+1)*bs)
        S=K[j0:j1,j0:j1].copy()
        for k0 in range(0,j0,16): S-=L[j0:j1,k0:k0+16]@L[j0:j1,k0:k0+16].T
        nb=j1-j0; M=np.tril(S).copy(); nt=(nb+7)//8; z=np.zeros(nb); yy=yt[j0:j1].copy(); Wt={}; Lt={}; fl={}
        for p in range(nt):
            a,b=8*p,min(nb,8*p+8)
            Lpp=np.linalg.cholesky(M[a:b,a:b]); Wpp=solve_triangular(Lpp,np.eye(b-a),lower=True)
            M[a:b,a:b]=Lpp; Wt[p]=Wpp; Lt[p]=Lpp; fl[p]=est(Lpp,Wpp)>tau; ntot+=1; nref+=fl[p]
            z[a:b]=Wpp@yy[a:b]
            if b<nb:
                T0=M[b:,a:b].copy(); X=T0@Wpp.T
                if fl[p]: X=X+(T0-X@Lpp.T)@Wpp.T
                M[b:,a:b]=X; M[b:,b:]-=np.tril(X@X.T); yy[b:]-=X@z[a:b]
        Ljj=np.tril(M); L[j0:j1,j0:j1]=Ljj
        quad+=float(z@z); logdet+=float(np.sum(np.log(np.diag(Ljj))))
        for I in range(J+1,T):
            i0,i1=I*bs,min(n,(I+1)*bs)
            Sij=K[i0:i1,j0:j1].copy()
            for k0 in range(0,j0,16): Sij-=L[i0:i1,k0:k0+16]@L[j0:j1,k0:k0+16].T
            X=np.zeros_like(Sij); Tm=Sij.copy()
            for c in range(nt):
                a,b=8*c,min(nb,8*c+8)
                X[:,a:b]=Tm[:,a:b]@Wt[c].T
                if fl[c]: X[:,a:b]+=(Tm[:,a:b]-X[:,a:b]@Lt[c].T)@Wt[c].T
                if b<nb: Tm[:,b:]-=X[:,a:b]@Ljj[b:,a:b].T
            L[i0:i1,j0:j1]=X; yt[i0:i1]-=X@z
    return -0.5*quad-0.5*(2*logdet+n*np.log(2*np.pi)), nref/ntot

rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 7)
rows=[]
ncase=0
while ncase<int(sys.argv[2]) if len(sys.argv)>2 else 40:
    n=int(rng.choice([rng.integers(20,130),rng.integers(130,700),rng.integers(700,1300)])); kind="gp" if rng.random()<.5 else "harmonic"
    bl=float(rng.uniform(0.3,3.0))*max(n,40)
    t,y,yerr,truth=synthetic_lightcurve(int(rng.integers(1<<30)),n,baseline=bl,kind=kind)
    if rng.random()<0.3: yerr=yerr*rng.uniform(0.5,20.0,size=n)
    th=o.sample_from_prior(4,seed=int(rng.integers(1<<30))); th[0]=truth
    for x in th:
        K=o.kernel_matrix(x,t,yerr)
        ev=np.linalg.eigvalsh(K); cond=ev[-1]/ev[0]
        if not (ev[0]>0): continue
        tl=E.ld(x,t,y,yerr); lap=o.lnlike_function(x,t,y,yerr)
        if not np.isfinite(lap): continue
        r={}
        for tau in (0,4,8,16,1e300):
            v,fr=lnlike_gpu(K,y,128,tau); r[tau]=(abs(v-tl)/abs(tl),fr)
        rows.append((cond,abs(lap-tl)/abs(tl),r)); ncase+=1
rows.sort(key=lambda q:q[0])
print("cond      lapack    always    tau=4 (frac refined)   tau=8       tau=16      never")
for c,l,r in rows:
    print("%.1e  %.1e   %.1e   %.1e (%.2f)   %.1e (%.2f)  %.1e (%.2f)  %.1e"%(c,l,r[0][0],r[4][0],r[4][1],r[8][0],r[8][1],r[16][0],r[16][1],r[1e300][0]))
import math
for tau in (4,8,16,1e300):
    ratios=[r[tau][0]/max(r[0][0],1e-17) for c,l,r in rows]
    worse=[ (r[tau][0], r[0][0], c) for c,l,r in rows if r[tau][0]>3*max(r[0][0],l,1e-15)]
    print("tau",tau,"median ratio to always-refine %.2f, max %.1f; cases >3x worse than max(always, lapack): %d"%(np.median(ratios),max(ratios),len(worse)), worse[:5])
