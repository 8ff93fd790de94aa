"""Randomized check of both Horn-Schunck SOR schedules (one-sweep: hs_sor_step.h, pipelined: hs_sor_pipe.h)
and of the two-columns-per-step step functions (hs_sor_pairs.h) against the sequential loop on the CPU: random shapes (3..99 x 3..129), TOL / maxiter, snapshot period,
prefetch distance, thread counts and adversaries.  CPU only.   python profiles/fuzz_hs_schedules.py [seconds]
Round 1: 16 780 cases (one-sweep + pipelined) and 8 419 cases (all three), 0 mismatches."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _hs_emu
from _hs_emu import run_seq, run_pipe_wave, run_wave, system
rs=np.random.RandomState(12345)
t0=time.time(); n=0; bad=0
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 420
while time.time()-t0 < budget:
    nx=int(rs.randint(3,100)); ny=int(rs.randint(3,130))
    ix,iy,rho,u,v,_=system(nx,ny,seed=int(rs.randint(1<<30)))
    tol=float(10**rs.uniform(-3.5,-0.3)) if rs.rand()<0.7 else 0.0
    maxiter=int(rs.randint(1,40))
    ref=run_seq(ix,iy,rho,u,v,7.0,tol,maxiter)
    K=int(rs.choice([1,2,3,5,8])); P=int(rs.randint(0,4)); nth=int(rs.choice([1,2,7,ny//2+1,ny,ny+3]))
    o,ph,la=int(rs.randint(3)),int(rs.randint(3)),int(rs.randint(3))
    g=run_pipe_wave(ix,iy,rho,u,v,7.0,tol,maxiter,K,P,nth,o,ph,la,seed=n)
    w=run_wave(ix,iy,rho,u,v,7.0,tol,maxiter,P,nth,o,ph,la,seed=n)
    q=run_pipe_wave(ix,iy,rho,u,v,7.0,tol,maxiter,K,P,nth,o,ph,la,seed=n,pairs=True)
    ok = g[2]==ref[2] and np.array_equal(g[0],ref[0]) and np.array_equal(g[1],ref[1]) and w[2]==ref[2] and np.array_equal(w[0],ref[0]) and q[2]==ref[2] and np.array_equal(q[0],ref[0]) and np.array_equal(q[1],ref[1])
    if not ok:
        bad+=1; print("MISMATCH",nx,ny,tol,maxiter,K,P,nth,o,ph,la,g[2],w[2],ref[2])
    n+=1
print("cases",n,"bad",bad)
