#!/usr/bin/env python
"""bench.py -- TV-L1 1080p frame-pairs/s on B200 (BASELINE.json metric), one JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
          --master-port P bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[2] -- a batch of synthetic 1920x1080 frame
pairs, default parameters (tau .25, lambda .15, theta .3, 5 scales x 5 warps, eps .01).  One step =
one pass of the solver over one batch of `--pairs` pairs per rank (default 256).  Pairs are
independent units, so ranks shard the batch with no data-path collective (weak scaling: every
rank gets its own `--pairs` pairs).

  value : whole-job frame-pairs/s, inputs already resident in HBM (tvl1_solve_batch_dev_f32)
  e2e   : same metric through the host-buffer C ABI (tvl1_solve_batch_f32): pinned host inputs,
          H2D and D2H inside the timed region
  roofline : the fused iteration kernel, algorithmic 64 B per pixel-iteration
             / CUDA-event time of its launches inside the timed region
  cpu_baseline : the reference's own CPU code (oracle/_ref, else the oracle port) on a bounded
             sample of the same workload, all host threads

--impl reference times that CPU implementation alone (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "TV-L1 1080p frame-pairs/sec"
UNIT = "frame-pairs/s"
ALGO_BYTES_PER_PIXEL_ITERATION = 64      # BASELINE.md section 3, SURVEY.md 8d
PARAMS = dict(tau=0.25, lam=0.15, theta=0.3, nscales=5, zfactor=0.5, warps=5, eps=0.01)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=256, help="frame pairs per rank per step")
    ap.add_argument("--e2e-pairs", type=int, default=256, help="frame pairs per rank per e2e step")
    ap.add_argument("--e2e-max-batch", type=int, default=16,
                    help="lock-step chunk of the host-buffer path (chunks alternate between two lanes)")
    ap.add_argument("--e2e-lanes", type=int, default=4,
                    help="lanes of the host-buffer call (4 measured best for lanes that do their own copies; 3 for TVL1_HOST_PIPE=1)")
    ap.add_argument("--nx", type=int, default=1920)
    ap.add_argument("--ny", type=int, default=1080)
    ap.add_argument("--max-batch", type=int, default=256, help="pairs advanced in lock-step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="batch1080p", choices=["batch1080p", "band4k", "band8k", "occ", "hs"],
                    help="batch1080p: BASELINE configs[2] (the headline); band4k / band8k: configs[3] / configs[4], "
                         "ONE pair split into row bands over the ranks (strong scaling); occ: the occlusion solver "
                         "(SURVEY 8f-3) on a batch of 640x480 frame triples per rank; hs: pyramidal Horn-Schunck (SURVEY 8f-4) on "
                         "296 x 1080p pairs, one GPU (profiles/bench_hs.py, same JSON contract)")
    ap.add_argument("--triples", type=int, default=148, help="occ workload: frame triples per rank per step")
    ap.add_argument("--no-row-band", action="store_true",
                    help="skip the row_band leg that the batch1080p workload appends when N > 1")
    ap.add_argument("--quick", action="store_true", help="skip the fp64 / sequence / copy-ceiling extras of e2e")
    return ap.parse_args()


BAND_CASES = {
    "band4k": dict(nx=3840, ny=2160, kw=dict(nscales=6, warps=10, eps=0.001), min_rows=1024,
                   name="BASELINE.json configs[3]: synthetic 3840x2160 pair, nscales=6 nwarps=10 eps=0.001"),
    "band8k": dict(nx=7680, ny=4320, kw=dict(nscales=5, warps=5, eps=0.01), min_rows=1024,
                   name="BASELINE.json configs[4]: synthetic 7680x4320 pair, default params"),
}


def config(args, n_gpus):
    return {
        "timed_region": "value: one tvl1_solve_batch_dev_f32 call per step with per-kernel-group CUDA events switched on "
                        "(profiling): the call returns when the step's events have completed, i.e. one host wait per step",
        "workload": "batch of %d synthetic %dx%d frame pairs per rank (BASELINE.json configs[2]), "
                    "5 scales x 5 warps, default params" % (args.pairs, args.nx, args.ny),
        "pairs_per_rank_per_step": args.pairs,
        "global_pairs_per_step": args.pairs * n_gpus,
        "nx": args.nx, "ny": args.ny,
        "params": PARAMS,
        "lockstep_batch": args.max_batch,
        "parallelism": "batch-sharded x%d, no collective" % n_gpus,
        "l2": "inputs (%.1f GB per rank) and solver state exceed the 126 MB L2; no explicit flush"
              % (args.pairs * 2 * args.nx * args.ny * 4 / 1e9),
    }


# ---- clocks -----------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, enabled=True):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        if not enabled:
            return
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            self.f.close()
            os.unlink(self.f.name)
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# ---- CPU reference leg ---------------------------------------------------------------------------
def cpu_solver():
    import numpy as np
    from oracle.loader import CpuTvl1, available
    kind = "reference" if available("reference", np.float64) else "port"
    return CpuTvl1(kind, np.float64), kind


def cpu_sample(args, budget_s, max_pairs=4):
    """Times the reference's CPU implementation on whole 1080p pairs of the same workload."""
    import numpy as np
    from optical_flow_1_b200 import synth
    cpu, kind = cpu_solver()
    cpu.set_threads(os.cpu_count() or 1)
    times = []
    t_all = time.perf_counter()
    for b in range(max_pairs):
        I0, I1 = synth.make_pair(args.nx, args.ny, seed=1234 + b)
        I0, I1 = I0.astype(np.float64), I1.astype(np.float64)
        t = time.perf_counter()
        cpu.multiscale(I0, I1, want_iters=False, **PARAMS)
        times.append(time.perf_counter() - t)
        if time.perf_counter() - t_all > budget_s:
            break
    return {
        "value": len(times) / sum(times), "unit": UNIT, "cores": cpu.max_threads(), "kind": kind,
        "sample": "%d of the workload's %dx%d pairs (seeds 1234..), Dual_TVL1_optic_flow_multiscale "
                  "in fp64, g++ -O3 -fopenmp, %d threads, %.1f s" % (
                      len(times), args.nx, args.ny, cpu.max_threads(), sum(times)),
    }


def run_reference(args, rank, world):
    if rank != 0:
        return
    import numpy as np
    from optical_flow_1_b200 import synth
    cpu, kind = cpu_solver()
    cpu.set_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1; use every host core
    nx, ny, prm, metric, cfg = args.nx, args.ny, PARAMS, METRIC, None
    if args.workload != "batch1080p":         # one pair of configs[3] / configs[4]: a step is the whole workload
        case = BAND_CASES[args.workload]
        nx, ny, prm = case["nx"], case["ny"], case["kw"]
        metric = "TV-L1 %dx%d frame-pairs/sec, one pair in row bands" % (nx, ny)
        cfg = {"workload": case["name"] + ", CPU reference (no split)", "params": prm}
    I0, I1 = synth.make_pair(nx, ny, seed=1234)
    I0, I1 = I0.astype(np.float64), I1.astype(np.float64)
    for _ in range(args.warmup):
        cpu.multiscale(I0, I1, want_iters=False, **prm)
    t = time.perf_counter()
    for _ in range(args.steps):
        cpu.multiscale(I0, I1, want_iters=False, **prm)
    dt = time.perf_counter() - t
    v = args.steps / dt
    sample = ("each step = 1 pair of the workload (%dx%d, seed 1234), fp64, g++ -O3 -fopenmp, %d threads"
              % (nx, ny, cpu.max_threads()))
    if cfg is not None:
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cpu.max_threads(), "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": dict(config(args, args.gpus), reference_arm_step="ONE pair of this workload per step "
                                            "(a bounded sample: the batch is 256 independent repetitions of this unit)"),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cpu.max_threads(), "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def pin_to_gpu_numa_node(index):
    """Best effort: run this rank (and allocate its pinned host buffers) on the CPUs next to its GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, mask in enumerate(words) for b in range(64) if (mask >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass



# ---- row bands: ONE image pair split over the ranks (BASELINE configs[3], [4]) -----
def band_measure(pkg, torch, dist, solver, case, rank, world, dev, steps=5, warmup=3, with_e2e=True):
    """Times the banded solve of one pair on `world` ranks (device-resident buffers, CUDA events on the
    solver's stream, barrier before every step, max over ranks), the ordinary single-GPU solve of the
    same pair on this rank, checks that the two agree bit for bit, and times the banded solve through
    the host-buffer C ABI (tvl1_band_solve_f32, pinned buffers)."""
    import numpy as np
    nx, ny, kw, min_rows = case["nx"], case["ny"], case["kw"], case["min_rows"]
    stream = torch.cuda.ExternalStream(solver.stream(), device=dev)
    I0, I1 = pkg.synth.make_batch_torch(1, nx, ny, seed=1234, device=dev)
    u1, u2 = torch.empty_like(I0), torch.empty_like(I0)
    b1, b2 = torch.empty_like(I0), torch.empty_like(I0)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        pkg.shard.barrier()

    def timed(fn, n):
        ms = []
        for _ in range(n):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            ms.append(pkg.shard.max_over_ranks(e0.elapsed_time(e1), dev))
        return ms

    solo = lambda: solver.solve_batch_device(I0.data_ptr(), I1.data_ptr(), u1.data_ptr(), u2.data_ptr(), 1, nx, ny,
                                             want_iters=True, **kw)
    for _ in range(warmup):
        it_solo, _ = solo()
    solo_ms = timed(solo, steps)
    solo_stats = solver.stats()
    out = {"case": case["name"], "nx": nx, "ny": ny, "params": kw, "n_gpus": world,
           "single_gpu_ms": min(solo_ms), "single_gpu_ms_mean": sum(solo_ms) / len(solo_ms),
           "iterations_per_level_fine_to_coarse": it_solo[0].sum(axis=1).tolist()[::-1]}
    if world < 2:
        out["pixel_iterations"] = solo_stats["pixel_iterations"]
        return out, (I0, I1, u1, u2)
    uid = [solver.band_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    solver.band_init(rank, world, uid[0])
    band = lambda: solver.band_solve_device(I0.data_ptr(), I1.data_ptr(), b1.data_ptr(), b2.data_ptr(), nx, ny,
                                            min_split_rows=min_rows, **kw)
    for _ in range(warmup):
        it_band, _ = band()
    band_ms = timed(band, steps)
    st = solver.stats()
    same = bool(torch.equal(u1, b1) and torch.equal(u2, b2) and np.array_equal(it_solo[0], it_band))
    same = pkg.shard.sum_over_ranks(0.0 if same else 1.0, dev) == 0.0
    out.update({
        "band_ms": min(band_ms), "band_ms_mean": sum(band_ms) / len(band_ms),
        "speedup_vs_single_gpu": min(solo_ms) / min(band_ms),
        "exchange": solver.band_exchange_mode(), "min_split_rows": min_rows,
        "band_rows_finest": [solver.band_rows(ny, r, world) for r in range(world)],
        "bit_identical_to_single_gpu_on_every_rank": same,
        "host_syncs_per_solve": st["host_syncs"],
        "halo": "4 rows of the six evolving planes pushed to each neighbour (NVLink stores fused into the "
                "iteration kernels) after every accepted block of <= 4 iterations; error sums through peer mailboxes",
    })
    if with_e2e:
        h = [torch.empty((ny, nx), dtype=torch.float32).pin_memory() for _ in range(4)]
        h[0].copy_(I0[0]); h[1].copy_(I1[0])
        torch.cuda.synchronize()
        prm = dict(kw)
        def host():
            solver.band_solve_host_ptr(h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), h[3].data_ptr(), nx, ny,
                                       min_split_rows=min_rows, **prm)
            return float(h[2][ny // 2, nx // 2])
        for _ in range(2):
            host()
        ts = []
        for _ in range(max(3, steps)):
            barrier()
            t = time.perf_counter()
            host()
            ts.append(pkg.shard.max_over_ranks(1e3 * (time.perf_counter() - t), dev))
        out["e2e_host_buffers_ms"] = min(ts)
        out["e2e_matches"] = bool(torch.equal(h[2], b1[0].cpu()) and torch.equal(h[3], b2[0].cpu()))
        out["e2e_h2d_bytes_per_rank"] = 2 * nx * ny * 4
        out["e2e_d2h_bytes_per_rank"] = 2 * nx * ny * 4
    out["pixel_iterations_this_rank"] = st["pixel_iterations"]
    return out, (I0, I1, b1, b2)


def run_band_workload(args, rank, local_rank, world):
    """--workload band4k | band8k: one step = one solve of ONE pair, split into row bands over the ranks
    (strong scaling: the work is fixed, N GPUs share it)."""
    import torch
    import torch.distributed as dist
    import optical_flow_1_b200 as pkg
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the solver has no CPU fallback)")
    case = BAND_CASES[args.workload]
    nx, ny = case["nx"], case["ny"]
    pin_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    solver = pkg.TVL1(device=local_rank, profiling=False)
    clocks = ClockSampler(local_rank, enabled=(rank == 0))
    res, bufs = band_measure(pkg, torch, dist, solver, case, rank, world, dev, steps=args.steps, warmup=args.warmup)
    clk = clocks.stop()
    ms = res["band_ms"] if world > 1 else res["single_gpu_ms"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # whole-solve algorithmic bytes: 64 B x pixel-iterations + 32 B x pixel-warps, all levels
    solver.set_profiling(True)
    solver.solve_batch_device(*[t.data_ptr() for t in bufs], 1, nx, ny, **case["kw"])
    st = solver.stats()
    algo = 64 * st["pixel_iterations"] + 32 * st["pixel_warps"]
    cpu_base = None
    if rank == 0 and not args.no_cpu_baseline:
        import numpy as np
        cpu, kind = cpu_solver()
        cpu.set_threads(os.cpu_count() or 1)
        I0, I1 = pkg.synth.make_pair(nx, ny, seed=1234)
        t = time.perf_counter()
        cpu.multiscale(I0.astype(np.float64), I1.astype(np.float64), want_iters=False, **case["kw"])
        dt = time.perf_counter() - t
        cpu_base = {"value": 1.0 / dt, "unit": UNIT, "cores": cpu.max_threads(), "kind": kind,
                    "sample": "the workload's one %dx%d pair, once, fp64, g++ -O3 -fopenmp, %d threads, %.1f s"
                              % (nx, ny, cpu.max_threads(), dt)}
    if rank == 0:
        e2e_ms = res.get("e2e_host_buffers_ms")
        line = {
            "metric": "TV-L1 %dx%d frame-pairs/sec, one pair in row bands" % (nx, ny), "value": 1e3 / ms, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": case["name"] + ", row-band split over %d GPU(s)" % world, "params": case["kw"],
                       "parallelism": "row bands x%d, peer-memory halo exchange fused into the iteration kernels" % world,
                       "l2": "state of the finest level (%.0f MB) exceeds the 126 MB L2" % (16 * nx * ny * 4 / 1e6)},
            "clocks": clk,
            "e2e": {"value": 1e3 / e2e_ms if e2e_ms else None, "unit": UNIT,
                    "h2d_bytes_per_step": 2 * nx * ny * 4 * world, "d2h_bytes_per_step": 2 * nx * ny * 4 * world,
                    "api": "tvl1_band_solve_f32 (every rank uploads the pair and receives the whole flow)"}
                   if e2e_ms else None,
            "gpu_launches": int(st["kernel_launches"]),
            "roofline": {"bound": "hbm", "kernel": "whole solve (all levels; iteration + warp kernels)",
                         "achieved": algo / (res["single_gpu_ms"] / 1e3) / 1e9 if world == 1 else
                                     algo / (ms / 1e3) / 1e9,
                         "peak": peak * world, "unit": "GB/s",
                         "frac": (algo / (ms / 1e3) / 1e9) / (peak * world), "traffic": None,
                         "note": "algorithmic bytes of the whole pair (64 B x pixel-iterations + 32 B x pixel-warps) "
                                 "over the solve time, against N x the measured HBM peak; the replicated coarse "
                                 "levels bound it (Amdahl)"},
            "cpu_baseline": cpu_base, "row_band": res,
        }
        print(json.dumps(line))
    solver.close()
    if world > 1:
        dist.destroy_process_group()


# ---- TV-L1 with occlusions ------------------------------------------------------------------
OCC_CASE = dict(nx=640, ny=480, kw=dict(lam=0.15, alpha=0.01, beta=0.15, theta=0.3, nscales=5, zfactor=0.5, warps=2, eps=0.01),
                name="batch of synthetic 640x480 frame triples (BASELINE.json configs[0]'s shape), CLI defaults of "
                     "src/tvl1occflow_constants.h:14-23 (5 scales by the CLI's size rule, 2 warps)")
OCC_UNIT = "frame-triples/s"
OCC_METRIC = "TV-L1 with occlusions, 640x480 frame-triples/sec"


def occ_cpu(args, budget_s, max_triples=3):
    """The reference's own CPU solver (oracle/_ref behind the zero-filling new[] shim, else the C
    restatement) on whole triples of the same workload, all host threads."""
    import numpy as np
    from oracle.loader import CpuOcc, occ_available
    from optical_flow_1_b200 import synth
    kind = "reference" if occ_available("reference", np.float64) else "port"
    cpu = CpuOcc(kind, np.float64)
    threads = os.cpu_count() or 1
    cpu.set_threads(threads)
    times = []
    t_all = time.perf_counter()
    for b in range(max_triples):
        T = synth.make_triple(OCC_CASE["nx"], OCC_CASE["ny"], seed=1234 + b)
        t = time.perf_counter()
        cpu.multiscale(T[0], T[1], T[2], None, **OCC_CASE["kw"])
        times.append(time.perf_counter() - t)
        if time.perf_counter() - t_all > budget_s:
            break
    return {"value": len(times) / sum(times), "unit": OCC_UNIT, "cores": threads, "kind": kind,
            "sample": "%d of the workload's triples (seeds 1234..), 7-plane Dual_TVL1_optic_flow_multiscale in fp64, "
                      "g++ -O3 -fopenmp, %d threads (its box relaxation is serial code), %.1f s"
                      % (len(times), threads, sum(times))}


def run_occ_reference(args, rank):
    if rank != 0:
        return
    b = occ_cpu(args, 1e9, max_triples=args.warmup)          # warm-up steps
    import numpy as np
    from oracle.loader import CpuOcc, occ_available
    from optical_flow_1_b200 import synth
    kind = "reference" if occ_available("reference", np.float64) else "port"
    cpu = CpuOcc(kind, np.float64)
    cpu.set_threads(os.cpu_count() or 1)
    T = synth.make_triple(OCC_CASE["nx"], OCC_CASE["ny"], seed=1234)
    t = time.perf_counter()
    for _ in range(args.steps):
        cpu.multiscale(T[0], T[1], T[2], None, **OCC_CASE["kw"])
    dt = time.perf_counter() - t
    v = args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": OCC_METRIC, "value": v, "unit": OCC_UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": OCC_CASE["name"], "params": OCC_CASE["kw"], "reference_arm_step": "ONE triple per step"},
        "cpu_baseline": dict(b, value=v, sample="each step = 1 triple of the workload (seed 1234)"),
        "e2e": {"value": v, "unit": OCC_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def run_occ_workload(args, rank, local_rank, world):
    """--workload occ: one step = one occ_solve_batch_dev_f64 call on `--triples` frame triples per rank
    (triples are independent: batch-sharded over the ranks, no collective)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import optical_flow_1_b200 as pkg
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the solver has no CPU fallback)")
    nx, ny, kw, B = OCC_CASE["nx"], OCC_CASE["ny"], OCC_CASE["kw"], args.triples
    pin_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    solver = pkg.TVL1Occ(device=local_rank, profiling=True, max_batch=args.triples)
    trip = [pkg.synth.make_triple(nx, ny, seed=1234 + rank * B + b) for b in range(B)]
    host = [np.stack([t[k] for t in trip]).astype(np.float64) for k in range(3)]
    dI = [torch.from_numpy(h).to(dev) for h in host]
    dO = [torch.empty_like(dI[0]) for _ in range(3)]
    ptrs = [t.data_ptr() for t in dI] + [0] + [t.data_ptr() for t in dO]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return solver.solve_batch_device(*ptrs, B, nx, ny, want_iters=True, **kw)

    for _ in range(args.warmup):
        step()
    clocks = ClockSampler(local_rank, enabled=(rank == 0))
    acc = None
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        iters, _ = step()
        st = solver.stats()
        acc = st if acc is None else {k: acc[k] + st[k] for k in st}
    barrier()
    wall = time.perf_counter() - t0
    clk = clocks.stop()
    ms = torch.tensor([1e3 * wall / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    # e2e: host buffers (pinned) through occ_solve_batch_f64, H2D of the three frames and D2H of flow + map inside
    hp = [torch.from_numpy(h).pin_memory() for h in host]
    ho = [torch.empty_like(hp[0]).pin_memory() for _ in range(3)]
    it_e = np.zeros((B, kw["nscales"], kw["warps"]), np.int32)
    prm = pkg.OccParams(kw["lam"], kw["alpha"], kw["beta"], kw["theta"], kw["nscales"], kw["zfactor"], kw["warps"], kw["eps"])
    import ctypes as C

    def e2e_step():
        rc = solver.lib.occ_solve_batch_f64(solver.ctx, C.c_int(B), C.c_void_p(hp[0].data_ptr()), C.c_void_p(hp[1].data_ptr()),
                                            C.c_void_p(hp[2].data_ptr()), None, C.c_void_p(ho[0].data_ptr()),
                                            C.c_void_p(ho[1].data_ptr()), C.c_void_p(ho[2].data_ptr()), C.c_int(nx),
                                            C.c_int(ny), C.byref(prm), it_e.ctypes.data_as(C.c_void_p), None)
        solver._ck(rc)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e_steps = max(1, min(args.steps, 3))
    for _ in range(e_steps):
        e2e_step()
    barrier()
    e_ms = torch.tensor([1e3 * (time.perf_counter() - t0) / e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
    e_ms = float(e_ms.item())
    same = bool(np.array_equal(it_e, iters) and torch.equal(ho[0].to(dev), dO[0]) and torch.equal(ho[2].to(dev), dO[2]))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    cpu_base = occ_cpu(args, args.cpu_seconds) if (rank == 0 and not args.no_cpu_baseline) else None
    if rank == 0:
        chi_bytes = 80.0 * acc["chi_pixel_iterations"]
        chi_gbps = chi_bytes / (acc["ms_chi"] / 1e3) / 1e9 if acc["ms_chi"] else None
        line = {
            "metric": OCC_METRIC, "value": B * world * 1e3 / ms, "unit": OCC_UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": OCC_CASE["name"], "params": kw, "triples_per_rank_per_step": B,
                       "global_triples_per_step": B * world, "parallelism": "batch-sharded x%d, no collective" % world,
                       "l2": "state of one step (%.1f GB per rank) exceeds the 126 MB L2; no explicit flush"
                             % (B * 65 * nx * ny * 8 / 1e9),
                       "timed_region": "value: wall clock around the steps (the outer loop reads one counter per outer "
                                       "iteration back), max over ranks"},
            "clocks": clk,
            "e2e": {"value": B * world * 1e3 / e_ms, "unit": OCC_UNIT, "ms_per_step": e_ms,
                    "h2d_bytes_per_step": 3 * B * nx * ny * 8 * world, "d2h_bytes_per_step": 3 * B * nx * ny * 8 * world,
                    "api": "occ_solve_batch_f64 (host pinned fp64 frames in, flow + occlusion map out)",
                    "matches_device_path": same},
            "gpu_launches": int(acc["kernel_launches"]),
            "roofline": {"bound": "hbm", "kernel": "k_occ_chi_eta + k_occ_chi_update: one primal-dual iteration of the "
                         "occlusion map (Solver_wrt_chi), 100 per outer iteration",
                         "achieved": chi_gbps, "peak": peak, "unit": "GB/s", "frac": chi_gbps / peak if chi_gbps else None,
                         "traffic": None, "algorithmic_bytes_per_pixel_iteration": 80,
                         "what": "fp64: read chi, eta1, eta2, g, F, G, beta div u; write chi, eta1, eta2",
                         "pixel_iterations": acc["chi_pixel_iterations"], "kernel_ms": acc["ms_chi"],
                         "kernel_share_of_step": acc["ms_chi"] / acc["ms_total"] if acc["ms_total"] else None},
            "box_relaxation": {"kernel": "k_occ_rof_gs (wavefront Gauss-Seidel, latency-bound: one CTA per plane)",
                               "cell_updates": acc["box_cell_updates"], "ms": acc["ms_box"],
                               "cell_updates_per_s": acc["box_cell_updates"] / (acc["ms_box"] / 1e3) if acc["ms_box"] else None,
                               "share_of_step": acc["ms_box"] / acc["ms_total"] if acc["ms_total"] else None},
            "device_ms_per_step_by_kernel_group": {k: acc[k] / args.steps for k in
                                                   ("ms_total", "ms_pyramid", "ms_warp", "ms_box", "ms_chi", "ms_other")},
            "outer_iterations_per_triple": acc["outer_iterations"] / args.steps / B,
            "host_syncs_per_step": acc["host_syncs"] / args.steps,
            "cpu_baseline": cpu_base,
        }
        print(json.dumps(line))
    solver.close()
    if world > 1:
        dist.destroy_process_group()


# ---- our arm -------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import optical_flow_1_b200 as pkg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the solver has no CPU fallback)")
    pin_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        pkg.shard.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return pkg.shard.max_over_ranks(x, dev)

    def sum_over_ranks(x):
        return pkg.shard.sum_over_ranks(x, dev)

    nx, ny, P = args.nx, args.ny, args.pairs
    solver = pkg.TVL1(device=local_rank, max_batch=args.max_batch, profiling=True)
    stream = torch.cuda.ExternalStream(solver.stream(), device=dev)

    # synthetic inputs generated straight into HBM; weak scaling: rank r holds pairs [r*P, (r+1)*P)
    # of a world*P-pair job, pair b made from seed 1234 + b
    first, _ = pkg.shard.shard_range(world * P, rank, world)
    I0, I1 = pkg.synth.make_batch_torch(P, nx, ny, seed=pkg.shard.pair_seed(1234, first), device=dev)
    u1 = torch.empty_like(I0)
    u2 = torch.empty_like(I0)
    torch.cuda.synchronize()

    def step_device():
        solver.solve_batch_device(I0.data_ptr(), I1.data_ptr(), u1.data_ptr(), u2.data_ptr(),
                                  P, nx, ny, **PARAMS)
        return solver.stats()

    for _ in range(args.warmup):
        step_device()

    barrier()
    clocks = ClockSampler(local_rank, enabled=(rank == 0))   # one poller per box: nvidia-smi queries perturb launches
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    acc = None
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        st = step_device()
        if acc is None:
            acc = st
        else:
            for k, v in st.items():
                acc[k] = [a + b for a, b in zip(acc[k], v)] if isinstance(v, list) else acc[k] + v
    e1.record(stream)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    dev_ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    ms = max_over_ranks(dev_ms)
    value = world * P * args.steps / (ms / 1e3)
    per_rank_ms = [dev_ms]
    if world > 1:
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = dev_ms
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        per_rank_ms = [float(x) for x in t.tolist()]

    # roofline of the fused iteration kernel over the timed region (this rank)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    it_ms = acc["iterate_ms"]
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    except (OSError, ValueError):
        pass
    per_level = []
    hbm_px = hbm_ms = chip_px = chip_ms = 0.0
    hbm_launches = 0
    blocked_mask = solver.blocked_levels()
    level_pixels, lx, ly = [], nx, ny                # pixels of every pyramid level (src/zoom.cpp:22-34)
    for l in range(PARAMS["nscales"]):
        level_pixels.append(lx * ly)
        lx, ly = solver.zoom_size(lx, ly, PARAMS["zfactor"])
    t1_level = None          # the streamed level where k_iterate_t1 runs alone (+ tail): the finest one
    blocked = []             # streamed levels served by the two-iterations-per-launch kernel
    # which levels the solver blocks follows the loop lengths of its previous solve (tvl1_ctx::t2_levels):
    # a level whose launches exceed its mean loop length x warps is running blocks + replays
    for l in range(PARAMS["nscales"]):
        lms, lpx, ll = acc["level_iterate_ms"][l], acc["level_pixel_iterations"][l], acc["level_iterate_launches"][l]
        # a level served by k_iterate_resident is ONE launch per warp step (the whole while loop on chip)
        on_chip = ll <= PARAMS["warps"] * args.steps
        gbs = ALGO_BYTES_PER_PIXEL_ITERATION * lpx / (lms / 1e3) / 1e9 if lms > 0 else None
        # the level's first launch where it is a two-iteration block from zero duals (k_iterate_t2): its own event span
        fb_ms = acc["level_first_block_ms"][l]
        fb_px = 2.0 * P * level_pixels[l] * args.steps if fb_ms > 0 else 0.0
        is_blocked = (not on_chip) and bool((blocked_mask >> l) & 1)
        kern = ("k_iterate_resident (on chip)" if on_chip else
                "k_iterate_t2 (two iterations per launch in registers) + k_iterate_t1 through HBM" if is_blocked else
                "k_iterate_t1 (+ k_iterate_tb in the tail launches) through HBM")
        per_level.append({"level": l, "kernel": kern, "launches": ll, "ms": lms, "GBps": gbs,
                          "frac_of_peak": gbs / peak if gbs else None,
                          "first_block": None if fb_ms <= 0 else {
                              "kernel": "k_iterate_t2 from zero duals: the first two iterations of every pair in one launch",
                              "ms": fb_ms, "pixel_iterations": fb_px,
                              "GBps": ALGO_BYTES_PER_PIXEL_ITERATION * fb_px / (fb_ms / 1e3) / 1e9}})
        if on_chip:
            chip_px += lpx; chip_ms += lms
        else:
            hbm_px += lpx; hbm_ms += lms; hbm_launches += ll
            if is_blocked:
                blocked.append(per_level[-1])
            elif t1_level is None:
                # k_iterate_t1 alone: the level's launches without its first block (a pair that stops after ONE
                # iteration has that block replayed, which this subtraction books as two iterations: none in this workload)
                t1_level = (lpx - fb_px, lms - fb_ms, ll - (args.steps if fb_ms > 0 else 0), l)
    # The roofline object is about the dominant HBM-streaming kernel ALONE: k_iterate_t1 on the finest level (the
    # level whose loops are too short to block: 1.9 iterations per warp step at default epsilon).  The level(s) that
    # run two iterations per launch move about half the algorithmic bytes and are reported beside it
    # (`blocked_levels`), like the on-chip levels, which move no HBM bytes per iteration.
    t_px, t_ms, t_launches, t_lv = t1_level if t1_level else (hbm_px, hbm_ms, hbm_launches, 0)
    achieved = ALGO_BYTES_PER_PIXEL_ITERATION * t_px / (t_ms / 1e3) / 1e9 if t_ms > 0 else None
    streamed = ALGO_BYTES_PER_PIXEL_ITERATION * hbm_px / (hbm_ms / 1e3) / 1e9 if hbm_ms > 0 else None
    all_levels = ALGO_BYTES_PER_PIXEL_ITERATION * acc["pixel_iterations"] / (it_ms / 1e3) / 1e9 if it_ms > 0 else None
    # the same kernels alone, measured now: 32 resident 1080p pairs, every pair iterating (tvl1_bench_iterate)
    kernel_only = None
    try:
        kernel_only = {}
        for name, mode, per in (("k_iterate_t1", "0", 1), ("k_iterate_t2", "2", 2)):
            os.environ["TVL1_BENCH_TB"] = mode
            ms_k = solver.bench_iterate(32, nx, ny, 20)
            g = ALGO_BYTES_PER_PIXEL_ITERATION * 32 * nx * ny * 20 * per / (ms_k / 1e3) / 1e9
            kernel_only[name] = {"iterations_per_launch": per, "ms_per_launch": ms_k / 20, "GBps_algorithmic": g,
                                 "frac_of_peak": g / peak}
        kernel_only["what"] = ("32 resident %dx%d pairs, 20 launches, every pair iterating; k_iterate_t2's launches include "
                               "the (empty) k_iterate_t1 launch of the same loop turn" % (nx, ny))
    except Exception as e:      # noqa: BLE001 -- diagnostics only
        kernel_only = {"error": str(e)}
    finally:
        os.environ.pop("TVL1_BENCH_TB", None)
    roofline = {
        "kernel": "k_iterate_t1 -- the fused primal-dual iteration (TH + div + u update + grad + p update + stop "
                  "test), one iteration per launch, on pyramid level %d (the finest: after the level's first block, "
                  "which runs two iterations per launch and is reported under per_level[].first_block, its loops are too "
                  "short for temporal blocking; the tail launches for the last few pairs use k_iterate_tb)" % t_lv,
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak if achieved else None,
        "frac_of_nominal_8000": achieved / 8000.0 if achieved else None,
        "traffic": traffic.get("dram_bytes_per_launch") if traffic else None,
        "traffic_source": traffic.get("source") if traffic else None,
        "traffic_launch": traffic.get("launch") if traffic else None,
        "traffic_algorithmic_bytes_same_launch": traffic.get("algorithmic_bytes_same_launch") if traffic else None,
        "peak_source": peak_src,
        "algorithmic_bytes_per_pixel_iteration": ALGO_BYTES_PER_PIXEL_ITERATION,
        "pixel_iterations": t_px, "launches": t_launches,
        "kernel_ms": t_ms, "kernel_share_of_step": t_ms / (dev_ms if dev_ms > 0 else 1),
        "blocked_levels": {"kernel": "k_iterate_t2 (+ k_iterate_t1 for single iterations and replays): two iterations per "
                                     "launch in registers, about 33 B of HBM traffic per pixel-iteration instead of 60",
                           "levels": blocked,
                           "share_of_step": sum(b["ms"] for b in blocked) / (dev_ms if dev_ms > 0 else 1)},
        "streamed_levels_GBps_equivalent": streamed,
        "streamed_levels_frac_of_peak": streamed / peak if streamed else None,
        "kernel_only": kernel_only,
        "on_chip_levels": {"kernel": "k_iterate_resident (cluster + DSMEM; 0 HBM bytes per iteration)",
                           "pixel_iterations": chip_px, "kernel_ms": chip_ms,
                           "algorithmic_GBps_equivalent": ALGO_BYTES_PER_PIXEL_ITERATION * chip_px / (chip_ms / 1e3) / 1e9 if chip_ms > 0 else None,
                           "share_of_step": chip_ms / (dev_ms if dev_ms > 0 else 1)},
        "all_iteration_launches_GBps_equivalent": all_levels,
        "per_level": per_level,
    }
    breakdown = {k: acc[k] / args.steps for k in ("total_ms", "iterate_ms", "warp_ms", "pyramid_ms",
                                                   "zoom_in_ms", "export_ms")}
    # secondary kernels against the same roofline (canonical bytes of BASELINE.md section 3)
    warp_gbs = 32 * acc["pixel_warps"] / (acc["warp_ms"] / 1e3) / 1e9 if acc["warp_ms"] > 0 else None
    # whole solve: canonical end-to-end bytes = 64 B x pixel-iterations + 32 B x pixel-warps
    # (+ pyramid, not counted) over the device time of the timed region
    whole_gbs = (64 * acc["pixel_iterations"] + 32 * acc["pixel_warps"]) / (dev_ms / 1e3) / 1e9 if dev_ms > 0 else None
    whole = {"algorithmic_GBps": whole_gbs, "frac_of_measured_peak": whole_gbs / peak if whole_gbs else None,
             "frac_of_nominal_8000": whole_gbs / 8000.0 if whole_gbs else None,
             "pixel_iterations_per_pair": acc["pixel_iterations"] / (P * args.steps),
             "note": "rank 0; 64 B x pixel-iterations + 32 B x pixel-warps, pyramid bytes not counted"}
    other = {"warp_precompute": {"algorithmic_bytes_per_pixel_warp": 32, "pixel_warps": acc["pixel_warps"],
                                 "achieved_GBps": warp_gbs, "frac": warp_gbs / peak if warp_gbs else None}}

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    solver.close()
    solver = pkg.TVL1(device=local_rank, max_batch=args.e2e_max_batch, profiling=False)
    solver.set_lanes(host_lanes=args.e2e_lanes)
    E = min(args.e2e_pairs, P)
    try:    # pinned host buffers of all ranks must stay a small part of the box's memory
        import psutil
        budget = 0.2 * psutil.virtual_memory().available / max(world, 1)
        while E > args.e2e_max_batch and 4 * E * nx * ny * 4 > budget:
            E //= 2
    except ImportError:
        pass
    hI0 = torch.empty((E, ny, nx), dtype=torch.float32).pin_memory()
    hI1 = torch.empty_like(hI0).pin_memory()
    hu1 = torch.empty_like(hI0).pin_memory()
    hu2 = torch.empty_like(hI0).pin_memory()
    hI0.copy_(I0[:E])
    hI1.copy_(I1[:E])
    torch.cuda.synchronize()

    def step_host():
        solver.solve_batch_host_ptr(hI0.data_ptr(), hI1.data_ptr(), hu1.data_ptr(), hu2.data_ptr(),
                                    E, nx, ny, dtype="float32", **PARAMS)
        return float(hu1[0, ny // 2, nx // 2])       # touch the result on the host

    for _ in range(max(1, min(args.warmup, 2))):
        step_host()
    barrier()
    e_steps = max(1, args.steps)
    t0 = time.perf_counter()
    for _ in range(e_steps):
        step_host()
    torch.cuda.synchronize()
    e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
    e2e_launches = solver.stats()["kernel_launches"]
    e2e = {"value": world * E * e_steps / (e_ms / 1e3), "unit": UNIT,
           "h2d_bytes_per_step": 2 * E * nx * ny * 4, "d2h_bytes_per_step": 2 * E * nx * ny * 4,
           "pairs_per_rank_per_step": E, "steps": e_steps, "ms_per_step": e_ms / e_steps,
           "api": "tvl1_solve_batch_f32 (host pinned fp32 in, host fp32 out)", "timer": "wall clock, max over ranks",
           "lockstep_batch": args.e2e_max_batch, "lanes": min(args.e2e_lanes, -(-E // args.e2e_max_batch))}
    # the device-resident result must equal the host-path result for the same pairs
    same = bool(torch.equal(hu1, u1[:E].cpu()) and torch.equal(hu2, u2[:E].cpu()))

    # copy-only ceiling of this box for the same bytes: every rank at once, H2D of the step's inputs and
    # D2H of a step's flows on two streams (full duplex), no kernels.  e2e cannot beat it.
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    dI = torch.empty((2, E, ny, nx), dtype=torch.float32, device=dev)

    def copy_only():
        with torch.cuda.stream(s_in):
            dI[0].copy_(hI0, non_blocking=True)
            dI[1].copy_(hI1, non_blocking=True)
        with torch.cuda.stream(s_out):
            hu1.copy_(u1[:E], non_blocking=True)      # (the same values they already hold)
            hu2.copy_(u2[:E], non_blocking=True)
        torch.cuda.synchronize()

    copy_only()
    barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        copy_only()
    c_ms = max_over_ranks(1e3 * (time.perf_counter() - t0)) / 2
    ceiling = world * E / (c_ms / 1e3)
    e2e["copy_ceiling"] = {"value": ceiling, "unit": UNIT, "ms_per_step": c_ms,
                           "GBps_each_direction_per_rank": 2 * E * nx * ny * 4 / (c_ms / 1e3) / 1e9,
                           "what": "same H2D + D2H bytes, all ranks simultaneously, pinned memory, two streams, no kernels"}
    e2e["frac_of_copy_ceiling"] = e2e["value"] / ceiling
    e2e["frac_of_device_rate"] = e2e["value"] / value
    del dI

    if not args.quick:
        # the same through the fp64 entry point -- the element type of the reference's own ABI
        # (ofpix_t = double): twice the PCIe bytes, narrowed / widened on the device
        E64 = max(4 * args.e2e_max_batch, E // 4)
        dI0 = hI0[:E64].double().pin_memory()
        dI1 = hI1[:E64].double().pin_memory()
        du1 = torch.empty_like(dI0).pin_memory()
        du2 = torch.empty_like(dI0).pin_memory()
        del hI0, hI1

        def step_host64():
            solver.solve_batch_host_ptr(dI0.data_ptr(), dI1.data_ptr(), du1.data_ptr(), du2.data_ptr(),
                                        E64, nx, ny, dtype="float64", **PARAMS)
            return float(du1[0, ny // 2, nx // 2])

        step_host64()
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            step_host64()
        torch.cuda.synchronize()
        e64_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
        same64 = bool(torch.equal(du1.float(), hu1[:E64]) and torch.equal(du2.float(), hu2[:E64]))
        e2e["fp64_abi"] = {"value": world * E64 * 2 / (e64_ms / 1e3), "unit": UNIT, "pairs_per_rank_per_step": E64,
                           "steps": 2, "h2d_bytes_per_step": 2 * E64 * nx * ny * 8, "d2h_bytes_per_step": 2 * E64 * nx * ny * 8,
                           "api": "tvl1_solve_batch_f64 (host pinned fp64 in, host fp64 out)", "matches_fp32_path": same64}

        # video form: F consecutive frames -> F-1 flows, each frame uploaded once (tvl1_solve_sequence_f32),
        # against the pairwise call on the same expanded pairs.  Frames alternate between the two images of
        # pair 0 (forward / backward flow), so every pair has realistic motion.
        del dI0, dI1, du1, du2
        S = E64
        hF = torch.empty((S + 1, ny, nx), dtype=torch.float32).pin_memory()
        hF[0::2] = I0[0].cpu()
        hF[1::2] = I1[0].cpu()
        sA = hF[:-1].clone().pin_memory()
        sB = hF[1:].clone().pin_memory()
        su1 = torch.empty((S, ny, nx), dtype=torch.float32).pin_memory()
        su2 = torch.empty_like(su1).pin_memory()
        pu1 = torch.empty_like(su1).pin_memory()
        pu2 = torch.empty_like(su1).pin_memory()

        def time_host(fn, reps=2):
            fn()
            barrier()
            t = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return max_over_ranks(1e3 * (time.perf_counter() - t)) / reps

        seq_ms = time_host(lambda: solver.solve_sequence_host_ptr(hF.data_ptr(), su1.data_ptr(), su2.data_ptr(),
                                                                  S + 1, nx, ny, **PARAMS))
        pair_ms = time_host(lambda: solver.solve_batch_host_ptr(sA.data_ptr(), sB.data_ptr(), pu1.data_ptr(),
                                                                pu2.data_ptr(), S, nx, ny, dtype="float32", **PARAMS))
        e2e["sequence"] = {"value": world * S / (seq_ms / 1e3), "unit": UNIT, "frames_per_rank_per_step": S + 1,
                           "h2d_bytes_per_step": (S + 1) * nx * ny * 4, "d2h_bytes_per_step": 2 * S * nx * ny * 4,
                           "api": "tvl1_solve_sequence_f32 (host pinned frames in, host fp32 flows out)",
                           "pairwise_same_pairs": world * S / (pair_ms / 1e3),
                           "matches_pairwise": bool(torch.equal(su1, pu1) and torch.equal(su2, pu2))}
        # the same video as 8-bit frames (tvl1_solve_sequence_u8): a quarter of the upload
        try:
            h8 = torch.clamp(torch.round(hF), 0, 255).to(torch.uint8).pin_memory()
            qu1 = torch.empty_like(su1).pin_memory()
            qu2 = torch.empty_like(su1).pin_memory()
            u8_ms = time_host(lambda: solver.solve_sequence_host_ptr(h8.data_ptr(), qu1.data_ptr(), qu2.data_ptr(),
                                                                     S + 1, nx, ny, dtype="uint8", **PARAMS))
            hW = h8.float().pin_memory()             # the same 8-bit frames widened on the host: the fp32 call's input
            wu1 = torch.empty_like(su1).pin_memory()
            wu2 = torch.empty_like(su1).pin_memory()
            w_ms = time_host(lambda: solver.solve_sequence_host_ptr(hW.data_ptr(), wu1.data_ptr(), wu2.data_ptr(),
                                                                    S + 1, nx, ny, **PARAMS))
            e2e["sequence_u8"] = {"value": world * S / (u8_ms / 1e3), "unit": UNIT, "frames_per_rank_per_step": S + 1,
                                  "h2d_bytes_per_step": (S + 1) * nx * ny, "d2h_bytes_per_step": 2 * S * nx * ny * 4,
                                  "api": "tvl1_solve_sequence_u8 (host pinned 8-bit frames in, host fp32 flows out); the frames "
                                         "are the video above rounded to 8 bits (quantisation changes the iteration counts, "
                                         "so compare with fp32_same_frames, not with the rows above)",
                                  "fp32_same_frames": world * S / (w_ms / 1e3),
                                  "matches_fp32_same_frames": bool(torch.equal(qu1, wu1) and torch.equal(qu2, wu2))}
            del h8, qu1, qu2, hW, wu1, wu2
        except Exception as e:          # an extra: never lose the headline line over it
            e2e["sequence_u8"] = {"error": repr(e)}
        del hF, sA, sB, su1, su2, pu1, pu2

    # ---- row bands (configs[3], configs[4]) on the same ranks: driver-visible at N > 1 ----
    row_band = None
    if world > 1 and not args.no_row_band:
        row_band = {}
        bs = pkg.TVL1(device=local_rank, profiling=False)
        for name in ("band4k", "band8k"):
            try:
                row_band[name], bufs = band_measure(pkg, torch, dist, bs, BAND_CASES[name], rank, world, dev, steps=3, warmup=2)
                del bufs
            except Exception as e:      # keep the headline line even if this leg fails
                row_band[name] = {"error": repr(e)}
            torch.cuda.empty_cache()
        bs.close()

    total_launches = sum_over_ranks(acc["kernel_launches"])
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_sample(args, args.cpu_seconds)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config(args, world), "clocks": clk, "e2e": e2e,
            "gpu_launches": int(total_launches), "roofline": roofline, "cpu_baseline": cpu_base,
            "timer": {"device_ms_rank0": dev_ms, "wall_ms_rank0": wall_ms,
                      "device_ms_per_rank": per_rank_ms},
            "host_syncs_per_step": acc["host_syncs"] / args.steps,
            "device_ms_per_step_by_kernel_group": breakdown,
            "other_kernels": other,
            "whole_solve": whole,
            "e2e_matches_device_path": same,
            "row_band": row_band,
        }
        print(json.dumps(line))
    solver.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload == "hs":
        # the Horn-Schunck bench line lives in profiles/bench_hs.py (one GPU; the reference's sweep is only
        # defined for one thread, so its CPU leg is the one-thread reference)
        if rank == 0:
            os.execv(sys.executable, [sys.executable, os.path.join(ROOT, "profiles", "bench_hs.py"), "--steps",
                                      str(args.steps), "--warmup", str(args.warmup)]
                     + (["--no-cpu"] if args.no_cpu_baseline else []))
        return
    if args.workload == "occ":
        run_occ_reference(args, rank) if args.impl == "reference" else run_occ_workload(args, rank, local_rank, world)
    elif args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload != "batch1080p":
        run_band_workload(args, rank, local_rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
