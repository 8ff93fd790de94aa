"""Horn-Schunck parity numbers for DESIGN section 10: GPU (fp32) against the one-thread CPU reference
(oracle/_ref when present, else the pinned oracle port) on the golden cases and on 640x480 with the
CLI's default parameters.  Prints one line per case: sweep counts equal?, mean / max |d| in px."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _cases
import optical_flow_1_b200 as pkg
from oracle.loader import CpuTvl1, available

cpu = CpuTvl1("reference" if available("reference", np.float64) else "port", np.float64)
gpu = pkg.HornSchunck(0)
cases = {k: (_cases.solver_inputs(c), c["kw"]) for k, c in _cases.HS_CASES.items()}
kw = dict(pkg.HS_DEFAULTS)
kw["nscales"] = pkg.hs_clamp_nscales(640, 480, kw["nscales"], kw["zfactor"])
cases["640x480_defaults"] = (_cases.synth.make_pair(640, 480, seed=1234), kw)
kw2 = dict(pkg.HS_DEFAULTS)
kw2["nscales"] = pkg.hs_clamp_nscales(1024, 436, kw2["nscales"], kw2["zfactor"])
cases["1024x436_defaults"] = (_cases.synth.make_pair(1024, 436, seed=1234), kw2)
for name, ((I1, I2), k) in cases.items():
    t0 = time.time()
    ru, rv, rit, rer = cpu.hs_multiscale(I1, I2, **k)
    t1 = time.time()
    u, v, it, er = gpu.horn_schunck_pyramidal(I1.astype(np.float32), I2.astype(np.float32), **k)
    t2 = time.time()
    d = np.concatenate([np.abs(u - ru).ravel(), np.abs(v - rv).ravel()])
    print("%-18s sweeps equal: %s (%d warp steps, %d sweeps)  mean|d| %.3g  max|d| %.3g  max|flow| %.2f  cpu(1 thread, %s) %.2fs  gpu %.2fs"
          % (name, bool(np.array_equal(it, rit)), it.size, int(it.sum()), d.mean(), d.max(), np.abs(ru).max(),
             cpu.kind, t1 - t0, t2 - t1))
