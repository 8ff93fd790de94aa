"""Sweep of the pinned host-buffer batch call (256 x 1080p through tvl1_solve_batch_f32): lanes that do their own copies
(TVL1_HOST_PIPE=0, the round-2 state) against the call-wide upload -> solve -> download pipeline with ramped chunk sizes."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import optical_flow_1_b200 as pkg
P, nx, ny = 256, 1920, 1080
I0, I1 = pkg.synth.make_batch_torch(P, nx, ny, seed=1234, device="cuda")
hI0 = torch.empty((P, ny, nx), dtype=torch.float32).pin_memory(); hI1 = torch.empty_like(hI0).pin_memory()
hu1 = torch.empty_like(hI0).pin_memory(); hu2 = torch.empty_like(hI0).pin_memory()
hI0.copy_(I0); hI1.copy_(I1); del I0, I1; torch.cuda.empty_cache()
ref = None
cfgs = [(16, 4, {"TVL1_HOST_PIPE": "0"}), (64, 3, {}), (64, 4, {}), (32, 4, {}), (32, 3, {}), (16, 4, {}), (64, 2, {}), (48, 3, {}),
        (64, 4, {"TVL1_CHUNKS": "8,8,16,32,64,64,32,16,8,8"}), (64, 4, {"TVL1_CHUNKS": "4,8,16,32,48,64,32,24,16,8,4"}),
        (96, 3, {"TVL1_CHUNKS": "8,16,32,96,64,24,16"}), (32, 6, {})]
if len(sys.argv) > 1:
    cfgs = eval(sys.argv[1])
for mb, lanes, env in cfgs:
    for k, v in env.items(): os.environ[k] = v
    g = pkg.TVL1(0, max_batch=mb, profiling=False)
    g.set_lanes(host_lanes=lanes)
    for k in env: del os.environ[k]
    best = 1e9
    try:
        times = []
        for rep in range(int(os.environ.get('REPS', '5'))):
            torch.cuda.synchronize(); t = time.perf_counter()
            g.solve_batch_host_ptr(hI0.data_ptr(), hI1.data_ptr(), hu1.data_ptr(), hu2.data_ptr(), P, nx, ny, dtype="float32")
            torch.cuda.synchronize(); dt = 1e3 * (time.perf_counter() - t)
            if rep: best = min(best, dt)
            times.append(round(dt, 1))
        if ref is None: ref = (hu1.clone(), hu2.clone())
        same = bool(torch.equal(ref[0], hu1) and torch.equal(ref[1], hu2))
        free, total = torch.cuda.mem_get_info()
        print("max_batch %3d lanes %d %-52s chunks %-40s: %.2f ms -> %.1f pairs/s  same=%s  dev mem used %.1f GB" % (
            mb, lanes, env, env.get("TVL1_CHUNKS") or (pkg.tvl1.plan_chunks(P, mb) if env.get("TVL1_HOST_PIPE") != "0" else "round-2 schedule"),
            best, P / best * 1e3, same, (total - free) / 2**30), flush=True)
        if os.environ.get('REPS'): print("    per call:", times, flush=True)
    except Exception as e:
        print("max_batch %3d lanes %d %s: FAILED %s" % (mb, lanes, env, e), flush=True)
    g.close(); del g; torch.cuda.empty_cache()
