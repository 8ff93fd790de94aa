"""Horn-Schunck parity table for DESIGN section 10.

Per case: the GPU path (fp32) against the one-thread CPU reference in fp64 (oracle/_ref when present,
else the pinned oracle port) -- warp steps whose sweep count differs, mean / max |d| in px -- and, with
--ref32, the same numbers for the reference's OWN float build against its double build: the yardstick for
what fp32 storage alone does to this solver on that input.

    python profiles/run_hs_parity.py            # on the GPU box
    python profiles/run_hs_parity.py --ref32 --no-gpu   # CPU only: the reference's float-vs-double table
"""
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

use_gpu = "--no-gpu" not in sys.argv
ref32 = "--ref32" in sys.argv
kind = "reference" if available("reference", np.float64) else "port"
cpu = CpuTvl1(kind, np.float64)
cpu32 = CpuTvl1(kind, np.float32) if ref32 else None
gpu = pkg.HornSchunck(0) if use_gpu else None


def defaults(nx, ny):
    kw = dict(pkg.HS_DEFAULTS)
    kw["nscales"] = pkg.hs_clamp_nscales(nx, ny, kw["nscales"], kw["zfactor"])
    return kw


cases = {k: (_cases.solver_inputs(c), c["kw"]) for k, c in _cases.HS_CASES.items()}
cases["640x480 CLI defaults"] = (_cases.synth.make_pair(640, 480, seed=1234), defaults(640, 480))
cases["1024x436 CLI defaults"] = (_cases.synth.make_pair(1024, 436, seed=1234), defaults(1024, 436))
cases["1920x1080 alpha=15 6x5 tol=1e-3 maxiter=60"] = (
    _cases.synth.make_pair(1920, 1080, seed=1234), dict(alpha=15.0, nscales=6, zfactor=0.5, warps=5, tol=1e-3, maxiter=60))
if "--defaults-1080p" in sys.argv:
    cases["1920x1080 CLI defaults"] = (_cases.synth.make_pair(1920, 1080, seed=1234), defaults(1920, 1080))
cache = os.path.join(ROOT, "profiles", "_cache", "hs_ref_1080p_a15.npz")


def diff(u, v, ru, rv, it, rit):
    d = np.concatenate([np.abs(u - ru).ravel(), np.abs(v - rv).ravel()])
    return "%2d of %2d warp steps differ, mean|d| %.3g max|d| %.3g" % (int((it != rit).sum()), it.size, d.mean(), d.max())


for name, ((I1, I2), k) in cases.items():
    t0 = time.time()
    if name.startswith("1920x1080 alpha") and os.path.exists(cache):
        z = np.load(cache)
        ru, rv, rit = z["u"], z["v"], z["iters"]
    else:
        ru, rv, rit, _ = cpu.hs_multiscale(I1, I2, **k)
    t1 = time.time()
    line = "%-44s max|flow| %6.2f  %5d sweeps  cpu %5.1fs" % (name, max(np.abs(ru).max(), np.abs(rv).max()),
                                                               int(rit.sum()), t1 - t0)
    if use_gpu:
        t2 = time.time()
        u, v, it, _ = gpu.horn_schunck_pyramidal(I1.astype(np.float32), I2.astype(np.float32), **k)
        line += "  gpu %5.2fs | GPU vs ref64: %s" % (time.time() - t2, diff(u, v, ru, rv, it, rit))
    if ref32:
        fu, fv, fit, _ = cpu32.hs_multiscale(I1.astype(np.float32), I2.astype(np.float32), **k)
        line += " | ref32 vs ref64: %s" % diff(fu, fv, ru, rv, fit, rit)
    print(line, flush=True)
