"""Device-resident 256 x 1080p pairs in lock-step chunks on 4 lanes (the execution shape of the host-buffer
path) under different limits of the temporally blocked kernel and tail settings: wall ms, best of 3.
    python profiles/run_chunks.py [max_batch lanes]..."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_1_b200 as pkg

P, nx, ny = 256, 1920, 1080
I0, I1 = pkg.synth.make_batch_torch(P, nx, ny, seed=1234, device="cuda")
u1, u2 = torch.empty_like(I0), torch.empty_like(I0)
shapes = [(16, 4), (32, 4), (64, 4)]
envs = [{}, {"TVL1_TB_MAX_MPIX": "20"}, {"TVL1_NO_TB": "1"}, {"TVL1_TAIL_PAIRS": "0"}, {"TVL1_NO_RESIDENT": "1"},
        {"TVL1_SLOT_CTAS": "8192"}]
for env in envs:
    for k in ("TVL1_TB_MAX_MPIX", "TVL1_NO_TB", "TVL1_TAIL_PAIRS", "TVL1_NO_RESIDENT", "TVL1_SLOT_CTAS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for mb, lanes in shapes:
        g = pkg.TVL1(0, max_batch=mb, profiling=False)
        g.set_lanes(host_lanes=4, dev_lanes=lanes)
        best = 1e9
        for rep in range(3):
            torch.cuda.synchronize()
            t = time.perf_counter()
            g.solve_batch_device(I0.data_ptr(), I1.data_ptr(), u1.data_ptr(), u2.data_ptr(), P, nx, ny)
            torch.cuda.synchronize()
            dt = 1e3 * (time.perf_counter() - t)
            if rep:
                best = min(best, dt)
        print("%-28s max_batch %3d lanes %d: %.2f ms -> %.1f pairs/s" % (env or "default", mb, lanes, best, P / best * 1e3), flush=True)
        g.close()
        del g
        torch.cuda.empty_cache()
