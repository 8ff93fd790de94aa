"""Where does a small lock-step chunk lose against the 256-pair batch?  Device-resident, ONE lane, B pairs per call:
ms per pair by kernel group and pyramid level (profiling events), with and without the shared-GPU kernel choice."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import optical_flow_1_b200 as pkg
nx, ny = 1920, 1080
Pmax = 256
I0, I1 = pkg.synth.make_batch_torch(Pmax, nx, ny, seed=1234, device="cuda")
u1 = torch.empty_like(I0); u2 = torch.empty_like(I0)
for B in [8, 16, 32, 64, 128, 256]:
    for notb in ([0, 1] if B < 256 else [0]):
        os.environ["TVL1_NO_TB"] = str(notb)
        g = pkg.TVL1(0, max_batch=B, profiling=True)
        g.set_lanes(dev_lanes=1)
        del os.environ["TVL1_NO_TB"]
        best = None
        for rep in range(4):
            torch.cuda.synchronize(); t = time.perf_counter()
            g.solve_batch_device(I0.data_ptr(), I1.data_ptr(), u1.data_ptr(), u2.data_ptr(), B, nx, ny)
            torch.cuda.synchronize(); dt = 1e3 * (time.perf_counter() - t)
            s = g.stats()
            if rep and (best is None or dt < best[0]): best = (dt, s)
        dt, s = best
        lv = " ".join("L%d %.3f(%d)" % (l, s["level_iterate_ms"][l] / B, s["level_iterate_launches"][l]) for l in range(5))
        print("B %3d no_tb %d: %.3f ms/pair wall | total %.3f iterate %.3f warp %.3f pyramid %.3f zoom_in %.3f export %.3f | %s" % (
            B, notb, dt / B, s["total_ms"] / B, s["iterate_ms"] / B, s["warp_ms"] / B, s["pyramid_ms"] / B, s["zoom_in_ms"] / B,
            s["export_ms"] / B, lv), flush=True)
        g.close(); del g
