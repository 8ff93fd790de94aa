"""256 x 1080p device-resident solve under different pair-slot targets (TVL1_SLOT_CTAS) and
temporal-blocking limits (TVL1_TB_MAX_MPIX): device time per level."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_1_b200 as pkg

P, nx, ny = 256, 1920, 1080
I0, I1 = pkg.synth.make_batch_torch(P, nx, ny, seed=1234, device="cuda")
u1, u2 = torch.empty_like(I0), torch.empty_like(I0)
ref = None
for slots, tbmax in [(8192, 192), (16384, 192), (32768, 192), (1 << 20, 192), (8192, 600), (32768, 600), (1 << 20, 600), (32768, 0)]:
    os.environ["TVL1_SLOT_CTAS"] = str(slots)
    os.environ["TVL1_TB_MAX_MPIX"] = str(tbmax)
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
    print("slot_ctas %7d tb_max %3d Mpx: total %.2f iterate %.2f warp %.2f | levels %s launches %s same=%s" % (
        slots, tbmax, best["total_ms"], best["iterate_ms"], best["warp_ms"],
        [round(x, 1) for x in best["level_iterate_ms"][:5]], best["level_iterate_launches"][:5], same), flush=True)
    g.close()
    del g
    torch.cuda.empty_cache()
