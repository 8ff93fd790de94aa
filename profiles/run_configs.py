"""BASELINE.json configs on one B200: GPU time (device-resident and through the host-buffer call),
iteration counts, and parity with the CPU reference (oracle/_ref when it travelled, else the port)
run on the box's host cores.  Writes a markdown table to stdout."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import optical_flow_1_b200 as pkg
from oracle.loader import CpuTvl1, available

CONFIGS = [
    ("configs[0] 640x480", 640, 480, dict(nscales=5, warps=5, eps=0.01)),
    ("configs[1] 1024x436", 1024, 436, dict(nscales=5, warps=5, eps=0.01)),
    ("configs[2] unit: one 1920x1080 pair", 1920, 1080, dict(nscales=5, warps=5, eps=0.01)),
    ("configs[3] 3840x2160", 3840, 2160, dict(nscales=6, warps=10, eps=0.001)),
    ("configs[4] 7680x4320", 7680, 4320, dict(nscales=5, warps=5, eps=0.01)),
]
only = [int(a) for a in sys.argv[1:]] or list(range(len(CONFIGS)))
kind = "reference" if available("reference", np.float64) else "port"
cpu = CpuTvl1(kind, np.float64)
cpu.set_threads(os.cpu_count() or 1)
g = pkg.TVL1(0, profiling=True)
print("| config | params | GPU device-resident ms | GPU host-buffer call ms (fp32 / fp64 drop-in) | CPU %s ms (%d threads) | speed-up (fp64 call) | iteration counts equal | mean / max |dflow| px | iterations per level (coarse->fine) | streaming iteration kernels (k_iterate_t1 + k_iterate_tb) GB/s at 64 B/px-iter |"
      % (kind, cpu.max_threads()))
print("|---|---|---|---|---|---|---|---|---|---|")
for idx in only:
    name, nx, ny, kw = CONFIGS[idx]
    I0, I1 = pkg.synth.make_pair(nx, ny, seed=1234)
    dI0, dI1 = torch.from_numpy(I0).cuda(), torch.from_numpy(I1).cuda()
    du1, du2 = torch.empty_like(dI0), torch.empty_like(dI0)
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize(); t = time.perf_counter()
        g.solve_batch_device(dI0.data_ptr(), dI1.data_ptr(), du1.data_ptr(), du2.data_ptr(), 1, nx, ny, **kw)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    st = g.stats()
    stream_ms = sum(st["level_iterate_ms"][l] for l in range(kw["nscales"]) if st["level_iterate_launches"][l] > kw["warps"])
    stream_px = sum(st["level_pixel_iterations"][l] for l in range(kw["nscales"]) if st["level_iterate_launches"][l] > kw["warps"])
    gbs = 64 * stream_px / (stream_ms / 1e3) / 1e9 if stream_ms > 0 else float("nan")
    t32 = t64 = 1e9
    I0d, I1d = I0.astype(np.float64), I1.astype(np.float64)
    for _ in range(3):
        t = time.perf_counter(); r32 = g.Dual_TVL1_optic_flow_multiscale(I0, I1, **kw); t32 = min(t32, time.perf_counter() - t)
        t = time.perf_counter(); r64 = g.Dual_TVL1_optic_flow_multiscale(I0d, I1d, **kw); t64 = min(t64, time.perf_counter() - t)
    t = time.perf_counter(); ref = cpu.multiscale(I0d, I1d, **kw); tc = time.perf_counter() - t
    d = np.concatenate([np.abs(r64[0] - ref[0]).ravel(), np.abs(r64[1] - ref[1]).ravel()])
    same = bool(np.array_equal(r64[2], ref[2]))
    ndiff = int((r64[2] != ref[2]).sum())
    print("| %s | %d scales x %d warps, eps %g | %.2f | %.2f / %.2f | %.0f | %.0fx | %s | %.2e / %.2e | %s | %.0f |"
          % (name, kw["nscales"], kw["warps"], kw["eps"], 1e3 * best, 1e3 * t32, 1e3 * t64, 1e3 * tc, tc / t64,
             "yes" if same else "no (%d of %d warps differ)" % (ndiff, r64[2].size), d.mean(), d.max(),
             r64[2].sum(axis=1).tolist(), gbs), flush=True)
