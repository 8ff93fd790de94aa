"""256 x 1080p device-resident solve under different settings of the two-phase lock-step loop
(TVL1_TAIL_PAIRS = active-pair count at which the loop switches to narrow launches, 0 = single-phase;
TVL1_TAIL_SLOT_CTAS = CTAs of a narrow launch; TVL1_TAIL_TB = temporal blocking in the narrow phase): device time per level, result must not change."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_1_b200 as pkg

P = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nx, ny = 1920, 1080
I0, I1 = pkg.synth.make_batch_torch(P, nx, ny, seed=1234, device="cuda")
u1, u2 = torch.empty_like(I0), torch.empty_like(I0)
ref = None
for tail, ctas, tb in [(0, 2048, 0), (16, 2048, 0), (16, 2048, 1), (32, 2048, 1), (64, 2048, 1), (32, 4096, 1), (100, 2048, 1)]:
    os.environ["TVL1_TAIL_PAIRS"] = str(tail)
    os.environ["TVL1_TAIL_SLOT_CTAS"] = str(ctas)
    os.environ["TVL1_TAIL_TB"] = str(tb)
    g = pkg.TVL1(0, max_batch=P, profiling=True)
    best = None
    for rep in range(3):
        g.solve_batch_device(I0.data_ptr(), I1.data_ptr(), u1.data_ptr(), u2.data_ptr(), P, nx, ny)
        st = g.stats()
        if rep and (best is None or st["total_ms"] < best["total_ms"]):
            best = st
    if ref is None:
        ref = (u1.clone(), u2.clone())
    same = bool(torch.equal(ref[0], u1) and torch.equal(ref[1], u2))
    gbs = [round(64 * px / (ms * 1e6), 0) if ms > 0 else None
           for px, ms in zip(best["level_pixel_iterations"][:5], best["level_iterate_ms"][:5])]
    print("tail_pairs %3d tail_ctas %5d tail_tb %d: total %.2f iterate %.2f warp %.2f | level ms %s GB/s %s launches %s same=%s" % (
        tail, ctas, tb, best["total_ms"], best["iterate_ms"], best["warp_ms"],
        [round(x, 1) for x in best["level_iterate_ms"][:5]], gbs, best["level_iterate_launches"][:5], same), flush=True)
    g.close()
    del g
    torch.cuda.empty_cache()
