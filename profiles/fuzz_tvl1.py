"""Randomised differential run of the GPU TV-L1 solver against the oracle (fp64 C restatement): random shapes
(incl. widths not divisible by 4), pyramid depths, zoom factors, warps and epsilons.  Prints one line per case and
a summary; the north_star bar is mean |d| <= 1e-3 px, max |d| <= 1e-2 px, equal iteration counts.
    python profiles/fuzz_tvl1.py [ncases] [seed]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optical_flow_1_b200 as pkg
from oracle.loader import CpuTvl1

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
gpu = pkg.TVL1(device=0)
cpu = CpuTvl1("port", np.float64)
cpu.set_threads(os.cpu_count() or 1)
bad = worst_mean = worst_max = 0
count_mismatch = 0
for c in range(ncases):
    nx, ny = int(rs.randint(48, 420)), int(rs.randint(40, 300))
    zf = float(rs.choice([0.5, 0.5, 0.6, 0.75]))
    nscales = int(rs.randint(1, 5))
    while nscales > 1 and min(nx, ny) * zf ** (nscales - 1) < 14:
        nscales -= 1
    kw = dict(tau=0.25, lam=float(rs.choice([0.1, 0.15, 0.3])), theta=float(rs.choice([0.2, 0.3, 0.5])), nscales=nscales,
              zfactor=zf, warps=int(rs.randint(1, 5)), eps=float(rs.choice([0.05, 0.01, 0.01, 0.004])))
    I0, I1 = pkg.synth.make_pair(nx, ny, seed=int(rs.randint(1, 10 ** 6)), scale=float(rs.uniform(0.2, 1.0)))
    g = gpu.Dual_TVL1_optic_flow_multiscale(I0, I1, **kw)
    r = cpu.multiscale(I0.astype(np.float64), I1.astype(np.float64), **kw)
    d = np.concatenate([np.abs(g[0] - r[0]).ravel(), np.abs(g[1] - r[1]).ravel()])
    same = bool(np.array_equal(g[2], r[2]))
    ok = d.mean() <= 1e-3 and d.max() <= 1e-2 and same
    bad += not ok
    count_mismatch += not same
    worst_mean, worst_max = max(worst_mean, d.mean()), max(worst_max, d.max())
    print("%3d %4dx%-4d zf %.2f scales %d warps %d eps %.3f lam %.2f theta %.1f: mean %.2e max %.2e iterations %s %s"
          % (c, nx, ny, zf, nscales, kw["warps"], kw["eps"], kw["lam"], kw["theta"], d.mean(), d.max(),
             "equal" if same else "DIFFER %s vs %s" % (g[2].ravel().tolist(), r[2].ravel().tolist()), "" if ok else "<-- outside the bar"),
          flush=True)
print("cases %d, outside the bar %d (iteration counts differ in %d), worst mean %.2e, worst max %.2e"
      % (ncases, bad, count_mismatch, worst_mean, worst_max))
