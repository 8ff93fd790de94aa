"""Single lane: where the temporally blocked kernel pays.  P x 1080p pairs in one lock-step batch (and one 4K /
8K pair) under different TVL1_TB_MAX_MPIX limits: wall ms per solve, best of 3."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_1_b200 as pkg

cases = [(256, 1920, 1080, {}), (64, 1920, 1080, {}), (16, 1920, 1080, {}), (4, 1920, 1080, {}), (1, 1920, 1080, {}),
         (1, 3840, 2160, dict(nscales=6, warps=10, eps=0.001)), (1, 7680, 4320, {})]
for P, nx, ny, kw in cases:
    I0, I1 = pkg.synth.make_batch_torch(P, nx, ny, seed=1234, device="cuda")
    u1, u2 = torch.empty_like(I0), torch.empty_like(I0)
    row = []
    for mpix in ("0", "5", "12", "40", "80", "192", "600"):
        os.environ["TVL1_TB_MAX_MPIX"] = mpix
        g = pkg.TVL1(0, max_batch=P, profiling=False)
        g.set_lanes(host_lanes=1, dev_lanes=1)
        best = 1e9
        for rep in range(3):
            torch.cuda.synchronize()
            t = time.perf_counter()
            g.solve_batch_device(I0.data_ptr(), I1.data_ptr(), u1.data_ptr(), u2.data_ptr(), P, nx, ny, **kw)
            torch.cuda.synchronize()
            dt = 1e3 * (time.perf_counter() - t)
            if rep:
                best = min(best, dt)
        row.append("%s:%.2f" % (mpix, best))
        g.close()
        del g
    print("%3d x %dx%d  ms by TB_MAX_MPIX  %s" % (P, nx, ny, "  ".join(row)), flush=True)
    del I0, I1, u1, u2
    torch.cuda.empty_cache()
