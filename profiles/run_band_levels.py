"""Per-level device time of one pair solved on one GPU and in row bands over the ranks (run under
torch.distributed.run): where the banded solve spends its time (replicated coarse levels vs split ones).
    python -m torch.distributed.run --nproc-per-node N ... profiles/run_band_levels.py [4k|8k] [min_split_rows]"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import optical_flow_1_b200 as pkg

which = sys.argv[1] if len(sys.argv) > 1 else "4k"
nx, ny, kw = (3840, 2160, dict(nscales=6, warps=10, eps=0.001)) if which == "4k" else (7680, 4320, dict(nscales=5, warps=5, eps=0.01))
min_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 512
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = pkg.TVL1(device=local, profiling=True)
I0, I1 = pkg.synth.make_batch_torch(1, nx, ny, seed=1234, device="cuda")
u1, u2 = torch.empty_like(I0), torch.empty_like(I0)
ptr = (I0.data_ptr(), I1.data_ptr(), u1.data_ptr(), u2.data_ptr())
keys = ("total_ms", "iterate_ms", "warp_ms", "pyramid_ms", "zoom_in_ms", "export_ms")
for _ in range(3):
    g.solve_batch_device(*ptr, 1, nx, ny, **kw)
st = g.stats()
out = {"case": which, "world": world, "single": {k: round(st[k], 3) for k in keys},
       "single_level_ms": [round(x, 3) for x in st["level_iterate_ms"][:kw["nscales"]]],
       "single_level_launches": st["level_iterate_launches"][:kw["nscales"]]}
if world > 1:
    uid = [g.band_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    g.band_init(rank, world, uid[0])
    for _ in range(3):
        torch.cuda.synchronize()
        dist.barrier()
        g.band_solve_device(*ptr, nx, ny, min_split_rows=min_rows, **kw)
    st = g.stats()
    out.update({"band": {k: round(st[k], 3) for k in keys},
                "band_level_ms": [round(x, 3) for x in st["level_iterate_ms"][:kw["nscales"]]],
                "band_level_launches": st["level_iterate_launches"][:kw["nscales"]], "min_split_rows": min_rows})
if rank == 0:
    print(json.dumps(out))
g.close()
if world > 1:
    dist.destroy_process_group()
