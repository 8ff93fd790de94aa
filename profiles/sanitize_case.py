"""Small end-to-end workload for compute-sanitizer: every kernel of the library runs at least once
(pyramid with zfactor 0.5 and 0.7, warp, streaming + cluster-resident iteration, zoom_in, export,
single-scale entry, batch of 3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optical_flow_1_b200 as pkg

g = pkg.TVL1(0)
I0, I1 = pkg.synth.make_pair(260, 180, seed=5, scale=0.5)
a = g.Dual_TVL1_optic_flow_multiscale(I0, I1, nscales=3, warps=2)
b = g.Dual_TVL1_optic_flow_multiscale(np.stack([I0, I1, I0]), np.stack([I1, I0, I0]), nscales=3, warps=2)
c = g.Dual_TVL1_optic_flow_multiscale(I0[:150, :201], I1[:150, :201], nscales=3, zfactor=0.7, warps=2)
d = g.Dual_TVL1_optic_flow(I0, I1, a[0], a[1], warps=1)
os.environ["TVL1_NO_RESIDENT"] = "1"
g2 = pkg.TVL1(0)
e = g2.Dual_TVL1_optic_flow_multiscale(I0, I1, nscales=3, warps=2)
print("ok", a[2].tolist(), e[2].tolist(), float(np.abs(a[0] - e[0]).max()))
