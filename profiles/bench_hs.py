"""Bench line of the pyramidal Horn-Schunck row, in the shape of
bench.py's contract: batch of synthetic 1920x1080 pairs, the CLI's default parameters
(src/horn_schunck_pyramidal_main.cpp:25-30, nscales clamped by :136-143).

  value        device-resident frame-pairs/s (inputs in HBM; CUDA events around the solve)
  e2e          the same through hs_solve_batch_f32 with pinned HOST buffers (H2D / D2H inside, wall clock)
  roofline     SOR kernel: 28 B x pixel-sweeps (counted on the device) / CUDA-event time of the SOR launches
               (incl. the two layout transposes per warp step) against the measured HBM copy peak
  cpu_baseline the unmodified reference (oracle/_ref) with ONE thread -- the only thread count for which
               its result is defined -- on one pair of the workload

    python profiles/bench_hs.py [--pairs 296] [--steps 1] [--warmup 1] [--e2e-pairs 296] [--no-cpu]
                                [--skip-device] [--e2e-warmup 1]

A solve takes about as long for 37 pairs as for 296 (one CTA per pair, up to two per SM), so the e2e
leg must carry the whole batch too: 4 lanes x 74 pairs.
"""
import argparse
import json
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import optical_flow_1_b200 as pkg

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=296)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--e2e-pairs", type=int, default=296)
ap.add_argument("--nx", type=int, default=1920)
ap.add_argument("--ny", type=int, default=1080)
ap.add_argument("--no-cpu", action="store_true")
ap.add_argument("--skip-device", action="store_true", help="only the e2e (host-buffer) leg")
ap.add_argument("--e2e-warmup", type=int, default=1)
ap.add_argument("--skip-e2e", action="store_true", help="only the device-resident leg")
ap.add_argument("--max-batch", type=int, default=148, help="device leg: lock-step batch (0: all pairs).  A launch lasts as "
                "long as its slowest pair (150 sweeps against a mean of 65), so two lock-step batches of 148 pairs on two "
                "lanes beat one of 296: 19.2 vs 15.2 pairs/s (profiles/r2o_hs_lanes.txt; 74 x 4 lanes 15.1, 37 x 8 12.7)")
ap.add_argument("--lanes", type=int, default=2, help="device leg: lanes that take lock-step batches concurrently")
args = ap.parse_args()
nx, ny, P = args.nx, args.ny, args.pairs
kw = dict(pkg.HS_DEFAULTS)
kw["nscales"] = pkg.hs_clamp_nscales(nx, ny, kw["nscales"], kw["zfactor"])

peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
peak = float(peaks.get("hbm_gbs", 6556.5))

I1, I2 = pkg.synth.make_batch_torch(P, nx, ny, seed=1234, device="cuda")
u, v = torch.empty_like(I1), torch.empty_like(I1)
MB = args.max_batch or P
g = pkg.HornSchunck(0, max_batch=MB, profiling=True)
g.set_lanes(host_lanes=4, dev_lanes=args.lanes)
stream = torch.cuda.ExternalStream(g.stream())


def solve():
    return g.hs_solve_batch_device(I1.data_ptr(), I2.data_ptr(), u.data_ptr(), v.data_ptr(), P, nx, ny,
                                   want_iters=True, **kw)


line = {}
it = None
if not args.skip_device:
    for _ in range(args.warmup):
        solve()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sor_ms = px = launches = sor_launches = 0
    import time as _time
    t0 = _time.perf_counter()
    with torch.cuda.stream(stream):
        e0.record()
        for _ in range(args.steps):
            it, er = solve()
            st = g.stats()
            sor_ms += st["iterate_ms"]; px += st["pixel_iterations"]
            launches += st["kernel_launches"]; sor_launches += st["iterate_launches"]
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if args.lanes > 1:      # the lanes run on their own streams: wall clock around the (synchronous) calls
        ms = 1e3 * (_time.perf_counter() - t0) / args.steps
    line = {
        "metric": "Horn-Schunck 1080p frame-pairs/sec", "value": P / (ms * 1e-3), "unit": "frame-pairs/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "batch of %d synthetic %dx%d frame pairs, pyramidal Horn-Schunck, CLI default parameters"
                               % (P, nx, ny), "params": kw, "lockstep_batch": MB, "lanes": args.lanes,
                   "l2": "inputs and solver state exceed the 126 MB L2; no explicit flush"},
        "gpu_launches": launches,
        "sweeps_per_warp_step_mean": float(it.mean()), "sweeps_per_warp_step_max": int(it.max()),
        "roofline": {"kernel": "k_hs_sor (+ k_hs_to_wave / k_hs_from_wave once per warp step)", "bound": "hbm",
                     "achieved": 28.0 * px / (sor_ms * 1e-3) / 1e9 if sor_ms else None, "peak": peak, "unit": "GB/s",
                     "frac": (28.0 * px / (sor_ms * 1e-3) / 1e9 / peak) if sor_ms else None,
                     "algorithmic_bytes_per_pixel_sweep": 28, "pixel_sweeps": px, "launches": sor_launches,
                     "kernel_ms": sor_ms, "kernel_share_of_step": sor_ms / (ms * args.steps),
                     "traffic": 17529608000, "traffic_source": "profiles/r1h_hs_sor_full.csv (148 pairs, finest level, "
                     "2 sweeps: 17.19 GB algorithmic)",
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6556.5"},
    }

if not args.skip_e2e:
    # e2e: pinned host buffers through the public batch call
    E = min(args.e2e_pairs, P)
    hI1 = torch.empty((E, ny, nx), dtype=torch.float32).pin_memory()
    hI2 = torch.empty_like(hI1).pin_memory()
    hu = torch.empty_like(hI1).pin_memory()
    hv = torch.empty_like(hI1).pin_memory()
    hI1.copy_(I1[:E]); hI2.copy_(I2[:E])
    torch.cuda.synchronize()
    h = pkg.HornSchunck(0, max_batch=max(1, (E + 3) // 4))     # 4 lanes
    import ctypes as C
    prm = h._hs_params(kw["alpha"], kw["nscales"], kw["zfactor"], kw["warps"], kw["tol"], kw["maxiter"])


    def host_solve():
        h._ck(h.lib.hs_solve_batch_f32(h.ctx, C.c_int(E), C.c_void_p(hI1.data_ptr()), C.c_void_p(hI2.data_ptr()),
                                       C.c_void_p(hu.data_ptr()), C.c_void_p(hv.data_ptr()), C.c_int(nx), C.c_int(ny),
                                       C.byref(prm), None, None))


    for _ in range(args.e2e_warmup):
        host_solve()
    t0 = time.perf_counter()
    host_solve()
    dt = time.perf_counter() - t0
    same = None if args.skip_device else bool(torch.equal(hu, u[:E].cpu()) and torch.equal(hv, v[:E].cpu()))
    line["e2e"] = {"value": E / dt, "unit": "frame-pairs/s", "pairs_per_step": E, "ms_per_step": dt * 1e3,
                   "h2d_bytes_per_step": 2 * E * nx * ny * 4, "d2h_bytes_per_step": 2 * E * nx * ny * 4,
                   "api": "hs_solve_batch_f32 (host pinned fp32 in, host fp32 out)", "matches_device_path": same,
                   "warmup": args.e2e_warmup, "lanes": 4, "lockstep_batch": max(1, (E + 3) // 4),
                   "finite": bool(torch.isfinite(hu).all() and torch.isfinite(hv).all())}

if not args.no_cpu and not args.skip_device:
    from oracle.loader import CpuTvl1, available
    kind = "reference" if available("reference", np.float64) else "port"
    cpu = CpuTvl1(kind, np.float64)
    a, b = I1[0].cpu().numpy().astype(np.float64), I2[0].cpu().numpy().astype(np.float64)   # exactly the GPU's input
    t0 = time.perf_counter()
    ru, rv, rit, _ = cpu.hs_multiscale(a, b, **kw)
    dt = time.perf_counter() - t0
    d = np.concatenate([np.abs(u[0].cpu().numpy() - ru).ravel(), np.abs(v[0].cpu().numpy() - rv).ravel()])
    line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "frame-pairs/s", "cores": 1, "kind": kind,
                            "sample": "pair 0 of the workload, fp64, g++ -O3, one thread (the reference's "
                                      "parallel sweep is a data race), %.1f s" % dt,
                            "sweeps_equal_to_gpu": bool(np.array_equal(rit, it[0])),
                            "mean_abs_diff_px": float(d.mean()), "max_abs_diff_px": float(d.max())}
print(json.dumps(line))
