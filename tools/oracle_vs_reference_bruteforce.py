"""Brute-force twin of tests/test_oracle_vs_reference_fuzz.py: random dense kernels and complete solves through the reference's own code (oracle/_ref) and the
restatement (strict build) for a given number of seconds; every result is compared with np.array_equal.  Development container only.
usage: python tools/oracle_vs_reference_bruteforce.py [seconds]   (round 2: 240 s -> 27 378 cases, 1200 s -> 122 420 cases, 0 mismatches)"""
import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp
import oracle as O
from oracle import ref as R
O.select_build("strict")
t0=time.time(); cnt=0; bad=0
rng=np.random.default_rng(12345)
while time.time()-t0 < (float(sys.argv[1]) if len(sys.argv) > 1 else 240):
    kind=rng.integers(0,4)
    if kind==0:
        m=int(rng.integers(2,64)); d=rng.standard_normal(m)*10.0**rng.integers(-3,4); e=rng.standard_normal(m-1)*10.0**rng.integers(-14,3)
        if rng.random()<0.3: e[rng.integers(0,m-1,max(1,(m-1)//3))]=0
        T=np.diag(d)+np.diag(e,1)+np.diag(e,-1)
        ok=all(np.array_equal(a,b) for a,b in zip(R.tridiag_eigen(T),O.tridiag_eigen(T)))
        sh=float(rng.standard_normal())
        ok&=all(np.array_equal(a,b) for a,b in zip(R.shifted_qr(T,sh,"tridiag"),O.shifted_qr(T,sh,"tridiag")))
    elif kind==1:
        m=int(rng.integers(3,50)); H=np.triu(rng.standard_normal((m,m)),-1)*10.0**rng.integers(-2,3)
        if rng.random()<0.3: H[np.arange(1,m),np.arange(0,m-1)]*=10.0**rng.integers(-16,0)
        ok=all(np.array_equal(a,b) for a,b in zip(R.hess_schur(H),O.hess_schur(H)))
        a,b=R.hess_eigen(H),O.hess_eigen(H); ok&=np.array_equal(a[0],b[0]) and np.array_equal(a[1],b[1])
        s_,t_=float(rng.standard_normal()),float(abs(rng.standard_normal())*3)
        ok&=all(np.array_equal(a,b) for a,b in zip(R.double_shift_qr(H,s_,t_),O.double_shift_qr(H,s_,t_)))
    elif kind==2:
        n=int(rng.integers(6,120)); M=rng.standard_normal((n,n)); M=M+M.T
        if rng.random()<0.3:
            r=max(1,n//5); B=rng.standard_normal((n,r)); M=B@B.T
        k=int(rng.integers(1,max(2,n//3))); m=int(min(n,max(k+2,2*k+1))); rule=[O.LargestMagn,O.LargestAlge,O.SmallestAlge,O.BothEnds,O.SmallestMagn][int(rng.integers(0,5))]
        fn=lambda x: M@x
        r=R.sym_eigs_userop(n,fn,k,m,selection=rule,maxit=150); o=O.sym_eigs_userop(n,fn,k,m,selection=rule,maxit=150)
        ok=(r.info,r.nconv,r.niter,r.nops)==(o.info,o.nconv,o.niter,o.nops) and np.array_equal(r.eigenvalues,o.eigenvalues) and (r.nconv==0 or np.array_equal(r.eigenvectors,o.eigenvectors))
    else:
        n=int(rng.integers(8,150)); A=sp.random(n,n,float(rng.uniform(0.03,0.5)),random_state=int(rng.integers(0,2**31)),format="csr")+sp.diags(rng.standard_normal(n)); A=sp.csr_matrix(A); A.sort_indices()
        k=int(rng.integers(1,max(2,n//4))); m=int(min(n,max(k+3,2*k+2))); rule=int(rng.integers(0,3)) if rng.random()<0.5 else int(rng.integers(4,7))
        r=R.gen_eigs(R.Compressed.from_scipy(A),k,m,rule,80); o=O.gen_eigs(O.Csr(n,A.indptr,A.indices,A.data,order="row",mode="gen"),k,m,rule,80)
        ok=(r.info,r.nconv,r.niter,r.nops)==(o.info,o.nconv,o.niter,o.nops) and np.array_equal(r.eigenvalues,o.eigenvalues) and (r.nconv==0 or np.array_equal(r.eigenvectors,o.eigenvectors))
    cnt+=1
    if not ok:
        bad+=1; print("MISMATCH kind",kind)
print("cases",cnt,"mismatches",bad)
