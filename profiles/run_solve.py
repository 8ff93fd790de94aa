"""Whole-solve driver for ncu launch lists: one device-resident batch of N synthetic pairs
(default 8 x 1920x1080), default parameters, solved twice (the first solve warms up)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_1_b200 as pkg

npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
ny = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
kw = dict(nscales=int(sys.argv[4]), warps=int(sys.argv[5]), eps=float(sys.argv[6])) if len(sys.argv) > 6 else {}
I0, I1 = pkg.synth.make_batch_torch(npairs, nx, ny, seed=1234, device="cuda")
u1, u2 = torch.empty_like(I0), torch.empty_like(I0)
torch.cuda.synchronize()
g = pkg.TVL1(0, max_batch=npairs, profiling=True)
for rep in range(2):
    it, er = g.solve_batch_device(I0.data_ptr(), I1.data_ptr(), u1.data_ptr(), u2.data_ptr(), npairs, nx, ny,
                                  want_iters=True, **kw)
    st = g.stats()
print("solve of %d pairs %dx%d: total %.2f ms  iterate %.2f  warp %.2f  pyramid %.2f  zoom_in %.2f  export %.2f  launches %d"
      % (npairs, nx, ny, st["total_ms"], st["iterate_ms"], st["warp_ms"], st["pyramid_ms"], st["zoom_in_ms"],
         st["export_ms"], st["kernel_launches"]))
print("iterations per (level, warp), pair 0:", it[0].tolist())
print("per level (0 = finest): launches", st["level_iterate_launches"][:it.shape[1]], "ms",
      [round(x, 2) for x in st["level_iterate_ms"][:it.shape[1]]])
