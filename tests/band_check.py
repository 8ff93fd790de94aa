"""Multi-GPU check of the row-band mode (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/band_check.py [nx ny nscales warps eps min_split_rows]

Every rank solves the pair twice -- alone (ordinary single-GPU solve) and together with the other
ranks (row bands, NCCL halo send/recv + error all-reduce per iteration) -- and the two results must
have identical iteration counts and flows within 1e-5 px (the pixel arithmetic is identical; only
the summation order of the error differs)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import optical_flow_1_b200 as pkg


def main():
    a = sys.argv[1:]
    nx, ny = (int(a[0]), int(a[1])) if len(a) >= 2 else (1024, 768)
    kw = dict(nscales=int(a[2]) if len(a) > 2 else 4, warps=int(a[3]) if len(a) > 3 else 3,
              eps=float(a[4]) if len(a) > 4 else 0.01)
    min_rows = int(a[5]) if len(a) > 5 else 256
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = pkg.TVL1(device=local)
    uid = [g.band_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    g.band_init(rank, world, uid[0])
    I0, I1 = pkg.synth.make_pair(nx, ny, seed=1234)

    solo = g.Dual_TVL1_optic_flow_multiscale(I0, I1, **kw)
    t = time.perf_counter()
    solo = g.Dual_TVL1_optic_flow_multiscale(I0, I1, **kw)
    t_solo = time.perf_counter() - t
    dist.barrier()
    band = g.band_solve(I0, I1, min_split_rows=min_rows, **kw)
    dist.barrier()
    t = time.perf_counter()
    band = g.band_solve(I0, I1, min_split_rows=min_rows, **kw)
    t_band = time.perf_counter() - t
    st = g.stats()

    # device-resident timings (no host copies): ordinary solve vs row bands
    dI0, dI1 = torch.from_numpy(I0).cuda(), torch.from_numpy(I1).cuda()
    du1, du2 = torch.empty_like(dI0), torch.empty_like(dI0)
    ptrs = (dI0.data_ptr(), dI1.data_ptr(), du1.data_ptr(), du2.data_ptr())

    def timed(fn, reps=3):
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return 1e3 * best

    dev_solo = timed(lambda: g.solve_batch_device(*ptrs, 1, nx, ny, **kw))
    big = max(min_rows, ny // 2 + 1)       # split the finest level only
    dev, same_top = {}, True
    for mode in ("peer", "nccl"):
        g.band_set_exchange(mode)
        eff = g.band_exchange_mode()
        dev[mode] = dict(mode_in_effect=eff,
                         all_levels_ge_min_rows=timed(lambda: g.band_solve_device(*ptrs, nx, ny, min_split_rows=min_rows, **kw)),
                         finest_level_only=timed(lambda: g.band_solve_device(*ptrs, nx, ny, min_split_rows=big, **kw)))
        it, _ = g.band_solve_device(*ptrs, nx, ny, min_split_rows=min_rows, **kw)
        same_top = same_top and bool(np.array_equal(it, solo[2]) and np.array_equal(du1.cpu().numpy(), solo[0])
                                     and np.array_equal(du2.cpu().numpy(), solo[1]))
    g.band_set_exchange("peer")

    d = max(np.abs(solo[0] - band[0]).max(), np.abs(solo[1] - band[1]).max())
    same_iters = bool(np.array_equal(solo[2], band[2]))
    rows = [g.band_rows(ny, r, world) for r in range(world)]
    res = dict(rank=rank, world=world, nx=nx, ny=ny, params=kw, min_split_rows=min_rows, band_rows=rows,
               same_iteration_counts=same_iters, max_abs_flow_diff=float(d),
               solo_ms=1e3 * t_solo, band_ms=1e3 * t_band, host_syncs=st["host_syncs"],
               device_resident_ms=dict(single_gpu=dev_solo, band=dev),
               both_exchange_modes_match_single_gpu=same_top, exchange_mode=g.band_exchange_mode(),
               iterations_per_level=band[2].sum(axis=1).tolist())
    ok = same_iters and d <= 1e-5 and same_top
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        res["all_ranks_ok"] = all(flags)
        print(json.dumps(res))
    g.close()
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
