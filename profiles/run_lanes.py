"""Device-resident batch of 256 x 1080p pairs under different (lock-step batch, device lanes)
splits: wall time of tvl1_solve_batch_dev_f32, best of 3."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_1_b200 as pkg

P = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nx, ny = 1920, 1080
I0, I1 = pkg.synth.make_batch_torch(P, nx, ny, seed=1234, device="cuda")
u1, u2 = torch.empty_like(I0), torch.empty_like(I0)
ref = None
for mb, lanes in [(256, 1), (128, 2), (64, 2), (64, 4), (32, 4), (128, 1)]:
    if mb > P:
        continue
    g = pkg.TVL1(0, max_batch=mb, profiling=False)
    g.set_lanes(host_lanes=3, dev_lanes=lanes)
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize()
        t = time.perf_counter()
        g.solve_batch_device(I0.data_ptr(), I1.data_ptr(), u1.data_ptr(), u2.data_ptr(), P, nx, ny)
        torch.cuda.synchronize()
        dt = 1e3 * (time.perf_counter() - t)
        if rep:
            best = min(best, dt)
    same = None
    if ref is None:
        ref = (u1.clone(), u2.clone())
    else:
        same = bool(torch.equal(ref[0], u1) and torch.equal(ref[1], u2))
    print("max_batch %3d lanes %d: %.2f ms  -> %.1f pairs/s  same=%s" % (mb, lanes, best, P / best * 1e3, same), flush=True)
    g.close()
    del g
    torch.cuda.empty_cache()
