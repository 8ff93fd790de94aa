"""What bounds the host-buffer batch call (round 2): 256 x 1080p pairs
  (1) device-resident under (lock-step batch, device lanes) splits -- the compute-only rate of the chunked execution;
  (2) through tvl1_solve_batch_f32 (pinned host buffers) under (lock-step batch, lanes, short first/last chunks).
Wall time per call, best of 3."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_1_b200 as pkg

P, nx, ny = 256, 1920, 1080
I0, I1 = pkg.synth.make_batch_torch(P, nx, ny, seed=1234, device="cuda")
u1, u2 = torch.empty_like(I0), torch.empty_like(I0)


def best_of(fn, reps=4):
    best = 1e9
    for rep in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        dt = 1e3 * (time.perf_counter() - t)
        if rep:
            best = min(best, dt)
    return best


for mb, lanes in [(256, 1), (64, 4), (32, 4), (16, 4)]:
    g = pkg.TVL1(0, max_batch=mb, profiling=False)
    g.set_lanes(host_lanes=4, dev_lanes=lanes)
    ms = best_of(lambda: g.solve_batch_device(I0.data_ptr(), I1.data_ptr(), u1.data_ptr(), u2.data_ptr(), P, nx, ny))
    print("device   max_batch %3d lanes %d: %.2f ms -> %.1f pairs/s" % (mb, lanes, ms, P / ms * 1e3), flush=True)
    g.close()
    del g
    torch.cuda.empty_cache()

hI0 = torch.empty((P, ny, nx), dtype=torch.float32).pin_memory()
hI1 = torch.empty_like(hI0).pin_memory()
hu1 = torch.empty_like(hI0).pin_memory()
hu2 = torch.empty_like(hI0).pin_memory()
hI0.copy_(I0)
hI1.copy_(I1)
ref = (u1.clone(), u2.clone())
del I0, I1
torch.cuda.empty_cache()
for mb, lanes, div in [(16, 4, 0), (16, 4, 4), (16, 4, 2), (32, 4, 0), (32, 4, 4), (32, 4, 8), (24, 4, 4), (64, 4, 8), (32, 3, 4)]:
    os.environ["TVL1_SHORT_DIV"] = str(div)
    g = pkg.TVL1(0, max_batch=mb, profiling=False)
    g.set_lanes(host_lanes=lanes)
    ms = best_of(lambda: g.solve_batch_host_ptr(hI0.data_ptr(), hI1.data_ptr(), hu1.data_ptr(), hu2.data_ptr(), P, nx, ny,
                                                dtype="float32"))
    same = bool(torch.equal(hu1.cuda(), ref[0]) and torch.equal(hu2.cuda(), ref[1]))
    print("host     max_batch %3d lanes %d short_div %d: %.2f ms -> %.1f pairs/s  same=%s"
          % (mb, lanes, div, ms, P / ms * 1e3, same), flush=True)
    g.close()
    del g
    torch.cuda.empty_cache()
