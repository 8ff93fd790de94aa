"""configs[3] (3840x2160, 6 scales x 10 warps, eps 0.001): where the GPU flow differs from the fp64
CPU reference -- histogram, location of the maximum, outer 3-px frame vs interior -- and the same for
the reference's own float build (what fp32 arithmetic costs on the CPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optical_flow_1_b200 as pkg
from oracle.loader import CpuTvl1, available

nx, ny = 3840, 2160
kw = dict(nscales=6, warps=10, eps=0.001)
if len(sys.argv) > 2:
    nx, ny = int(sys.argv[1]), int(sys.argv[2])
I0, I1 = pkg.synth.make_pair(nx, ny, seed=1234)
kind = "reference" if available("reference", np.float64) else "port"
c64 = CpuTvl1(kind, np.float64); c64.set_threads(os.cpu_count() or 1)
ref = c64.multiscale(I0.astype(np.float64), I1.astype(np.float64), **kw)
g = pkg.TVL1(0)
out = g.Dual_TVL1_optic_flow_multiscale(I0, I1, **kw)


def report(name, u1, u2):
    d = np.maximum(np.abs(u1 - ref[0]), np.abs(u2 - ref[1]))
    i, j = np.unravel_index(np.argmax(d), d.shape)
    inner = d[3:-3, 3:-3]
    print("%s: mean %.3e max %.3e at (row %d, col %d); interior (3-px frame removed) max %.3e; pixels > 1e-3: %d, > 1e-2: %d of %d"
          % (name, d.mean(), d.max(), i, j, inner.max(), int((d > 1e-3).sum()), int((d > 1e-2).sum()), d.size))
    ys, xs = np.nonzero(d > 1e-2)
    if len(ys):
        print("   rows of > 1e-2 pixels: %d..%d, cols %d..%d; flow there: ref u=(%.3f, %.3f)" % (ys.min(), ys.max(), xs.min(), xs.max(), ref[0][i, j], ref[1][i, j]))


print("iteration counts equal:", bool(np.array_equal(out[2], ref[2])))
report("GPU fp32 vs CPU fp64", out[0].astype(np.float64), out[1].astype(np.float64))
if available("reference", np.float32):
    c32 = CpuTvl1("reference", np.float32); c32.set_threads(os.cpu_count() or 1)
    r32 = c32.multiscale(I0, I1, **kw)
    print("CPU float build iteration counts equal to fp64:", bool(np.array_equal(r32[2], ref[2])))
    report("CPU fp32 (reference, ofpix_t=float) vs CPU fp64", r32[0].astype(np.float64), r32[1].astype(np.float64))
