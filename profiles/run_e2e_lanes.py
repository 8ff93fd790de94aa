import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import optical_flow_1_b200 as pkg
P, nx, ny = 256, 1920, 1080
I0, I1 = pkg.synth.make_batch_torch(P, nx, ny, seed=1234, device="cuda")
hI0 = torch.empty((P, ny, nx), dtype=torch.float32).pin_memory(); hI1 = torch.empty_like(hI0).pin_memory()
hu1 = torch.empty_like(hI0).pin_memory(); hu2 = torch.empty_like(hI0).pin_memory()
hI0.copy_(I0); hI1.copy_(I1); del I0, I1; torch.cuda.empty_cache()
for mb, lanes, div in [(16, 4, 2), (16, 6, 2), (16, 8, 2), (32, 6, 2), (32, 8, 2), (8, 8, 0), (12, 6, 0)]:
    os.environ["TVL1_SHORT_DIV"] = str(div)
    g = pkg.TVL1(0, max_batch=mb, profiling=False)
    g.set_lanes(host_lanes=lanes)
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize(); t = time.perf_counter()
        g.solve_batch_host_ptr(hI0.data_ptr(), hI1.data_ptr(), hu1.data_ptr(), hu2.data_ptr(), P, nx, ny, dtype="float32")
        torch.cuda.synchronize(); dt = 1e3 * (time.perf_counter() - t)
        if rep: best = min(best, dt)
    print("host max_batch %3d lanes %d short_div %d: %.2f ms -> %.1f pairs/s" % (mb, lanes, div, best, P / best * 1e3), flush=True)
    g.close(); del g; torch.cuda.empty_cache()
