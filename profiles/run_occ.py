"""Driver for ncu captures of the occlusion solver: B triples of nx x ny, default parameters, one solve.
    python profiles/run_occ.py [B nx ny nscales]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import optical_flow_1_b200 as pkg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 640
ny = int(sys.argv[3]) if len(sys.argv) > 3 else 480
ns = int(sys.argv[4]) if len(sys.argv) > 4 else 2
trip = [pkg.synth.make_triple(nx, ny, seed=1234 + b) for b in range(B)]
dI = [torch.from_numpy(np.stack([t[k] for t in trip]).astype(np.float64)).cuda() for k in range(3)]
dO = [torch.empty_like(dI[0]) for _ in range(3)]
g = pkg.TVL1Occ(0, profiling=True, max_batch=B)
it, _ = g.solve_batch_device(dI[0].data_ptr(), dI[1].data_ptr(), dI[2].data_ptr(), 0, dO[0].data_ptr(), dO[1].data_ptr(),
                             dO[2].data_ptr(), B, nx, ny, want_iters=True, nscales=ns, warps=1)
print(g.stats(), it.sum())
