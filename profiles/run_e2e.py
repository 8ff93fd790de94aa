"""Host-buffer (pinned) batch of 256 x 1080p pairs through tvl1_solve_batch_f32 under different
(lock-step batch, lanes) splits: wall time per call, best of 3."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_1_b200 as pkg

P, nx, ny = 256, 1920, 1080
I0, I1 = pkg.synth.make_batch_torch(P, nx, ny, seed=1234, device="cuda")
hI0 = torch.empty((P, ny, nx), dtype=torch.float32).pin_memory()
hI1 = torch.empty_like(hI0).pin_memory()
hu1 = torch.empty_like(hI0).pin_memory()
hu2 = torch.empty_like(hI0).pin_memory()
hI0.copy_(I0)
hI1.copy_(I1)
del I0, I1
torch.cuda.empty_cache()
cfgs = [(32, 4), (16, 4), (24, 4), (48, 4), (64, 4), (32, 3), (32, 2), (64, 2)]
if len(sys.argv) > 1:
    cfgs = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for mb, lanes in cfgs:
    g = pkg.TVL1(0, max_batch=mb, profiling=False)
    g.set_lanes(host_lanes=lanes)
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize()
        t = time.perf_counter()
        g.solve_batch_host_ptr(hI0.data_ptr(), hI1.data_ptr(), hu1.data_ptr(), hu2.data_ptr(), P, nx, ny, dtype="float32")
        dt = 1e3 * (time.perf_counter() - t)
        if rep:
            best = min(best, dt)
    print("max_batch %3d lanes %d: %.2f ms -> %.1f pairs/s" % (mb, lanes, best, P / best * 1e3), flush=True)
    g.close()
    del g
    torch.cuda.empty_cache()
