#!/usr/bin/env python3
"""How well is the CG iteration count of the reference's Poisson solve (benamou_brenier.py:85, scipy cg, rtol 1e-6)
defined?  CPU only; needs /root/reference for the operators (like tests/golden/make_golden.py).  Runs scipy's
recurrence on the first stepA system of a 97x146x4 pair three times: np.dot, a dot product summed in another order,
and the single-reduction (Chronopoulos-Gear) arrangement of the CUDA kernel, and prints the relative difference of
||r_k||^2 along the way.  Result (profiles/r2_cg_count_sensitivity.txt): 1e-5..1e-4 after a few hundred iterations for
BOTH changes, against a decay of 3-9 % per iteration -- a solve whose final residual lands within ~1e-4 of the
threshold ends one iteration earlier or later, i.e. about 1 solve in 300, whatever the implementation."""
import sys, numpy as np, math
sys.path.insert(0,'/root/reference'); sys.path.insert(0,'/root/repo/optical-flow-optimal-transport_b200')
import operators, scipy.sparse as sp
from foto_b200 import synth
h,w,Nt=97,146,4
f0,f1=synth.make_pair(h,w,seed=3)
N=Nt*h*w;P=h*w
L=operators.laplacian_st(Nt,w,h,1,1,1,bc='N'); D=operators.div_st(Nt,w,h,1,1,1,bc='N')
A=(-1.0*L+1e-3*sp.eye(N)).tocsr()
mu=np.zeros(3*N)
for n in range(Nt): mu[n*P:(n+1)*P]=(1-n/(Nt-1))*f0+n/(Nt-1)*f1
F=D@mu; F[:P]-=f0-mu[:P]; F[(Nt-1)*P:N]+=f1-mu[(Nt-1)*P:N]
def cg(dot, variant='textbook', maxit=1000):
    x=np.zeros(N); r=F.copy(); hist=[]
    atol=1e-6*math.sqrt(dot(F,F))
    if variant=='textbook':
        rho_prev=None; p=None
        for k in range(maxit):
            rho=dot(r,r); hist.append(rho)
            if math.sqrt(rho)<atol: return k,hist
            p = r.copy() if p is None else r+(rho/rho_prev)*p
            q=A@p; alpha=rho/dot(p,q); x+=alpha*p; r-=alpha*q; rho_prev=rho
    else:  # Chronopoulos-Gear
        p=np.zeros(N); s=np.zeros(N); gam_old=None; alpha_old=None
        for k in range(maxit):
            wv=A@r; gam=dot(r,r); delt=dot(r,wv); hist.append(gam)
            if math.sqrt(gam)<atol: return k,hist
            if gam_old is None: beta=0.0; alpha=gam/delt
            else: beta=gam/gam_old; alpha=gam/(delt-beta*gam/alpha_old)
            p=r+beta*p; s=wv+beta*s; x+=alpha*p; r-=alpha*s; gam_old=gam; alpha_old=alpha
    return maxit,hist
d1=lambda a,b: float(np.dot(a,b))
d2=lambda a,b: float(np.sum((a*b).reshape(-1,8).sum(axis=1)))   # different summation order
k1,h1=cg(d1); k2,h2=cg(d2); k3,h3=cg(d1,'cg')
print('iters',k1,k2,k3)
n=min(len(h1),len(h2),len(h3))
for k in [10,100,200,300,400,500,n-1]:
    print(k, 'rel diff textbook(dot order) %.2e  CG-vs-textbook %.2e   decay/iter %.4f'%(abs(h1[k]-h2[k])/h1[k], abs(h1[k]-h3[k])/h1[k], math.sqrt(h1[k]/h1[k-1])))
