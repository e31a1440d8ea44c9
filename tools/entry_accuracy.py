"""Entry-level accuracy of the two covariance synthesis modes of gp_kernel.cuh against a long-double evaluation of the model,
for the one row of the bench launch whose cond(K) is 7.3e9 (row 974): george's operation order (GPROT_EXACT_KERNEL, what the
oracle computes too) and the fast mode (phases reduced once per point, sin of the difference from the addition theorem, one
exp).  NumPy emulation with the fused multiply-adds done in long double.  Both modes are 1e-14 accurate at worst and 2e-16 in
the median: at eps * cond = 1.6e-6 the three fp64 answers for that row (GPU fast 6.8e-7, GPU exact 7.8e-8, LAPACK 6e-8 ... 2.7e-7
depending on the host) are different roundings of equally good entries.  Test infrastructure (imports oracle/)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import gp_oracle as o
from gprotation_b200.synth import synthetic_lightcurve
t,y,yerr,_=synthetic_lightcurve(0,2048,kind="gp"); th=o.sample_from_prior(1024,seed=1000)
x=th[974]; A,L,G,S,P=np.exp(x)
print("theta",x,"A %.3g L %.3g G %.3g S %.3g P %.3g"%(A,L,G,S,P))
ld=np.longdouble
n=400                      # a corner of the matrix is enough for entry statistics
ti=t[:n]
# truth in long double (inputs are the doubles)
d=ti[:,None].astype(ld)-ti[None,:].astype(ld)
Kt=ld(A)*np.exp(-ld(0.5)*d*d/ld(L) - ld(G)*np.sin(ld(np.pi)*d/ld(P))**2)
# george order in double
dd=ti[:,None]-ti[None,:]
r2=(dd*dd)/L; s=np.sin((np.pi*dd)/P); Ke=(A*np.exp(-0.5*r2))*np.exp((-G*s)*s)
# fast mode in double: phases reduced with an emulated fma (long double), sincospi via numpy on the reduced phase
hi=ti/P; lo=((ti.astype(ld)-hi.astype(ld)*ld(P)).astype(np.float64))/P
ph=(hi-np.rint(hi))+lo
si=np.sin(np.pi*ph.astype(ld)).astype(np.float64); ci=np.cos(np.pi*ph.astype(ld)).astype(np.float64)   # sincospi is correctly rounded-ish
sn=(si[:,None].astype(ld)*ci[None,:].astype(ld) - (ci[:,None]*si[None,:]).astype(ld)).astype(np.float64)        # fma(si,cj,-(ci*sj))
hl=min(0.5/L,1.7e308)
xarg=((-G*sn).astype(ld)*sn.astype(ld) + (-(dd*dd)*hl).astype(ld)).astype(np.float64)                  # fma
Kf=A*np.exp(xarg)                                                                                   # exp_neg <= 1 ulp: use libm here
def stats(K,name):
    rel=np.abs((K.astype(ld)-Kt)/Kt).astype(np.float64)
    m=np.abs(Kt)>1e-300
    print(name,"relative entry error vs long double: median %.2e  99%% %.2e  max %.2e   (eps=1.1e-16)"%(np.median(rel[m]),np.quantile(rel[m],.99),rel[m].max()))
    near=np.abs(dd)<5*np.sqrt(L) 
    print("     entries with |d| < 5 sqrt(L): max %.2e"%rel[m&near].max(), " weighted by K/A: max %.2e"%(rel*np.abs(Kt.astype(np.float64))/A)[m].max())
stats(Ke,"george order"); stats(Kf,"fast mode   ")
