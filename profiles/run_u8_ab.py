import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import optical_flow_1_b200 as pkg
S, nx, ny = 64, 1920, 1080
I0, I1 = pkg.synth.make_batch_torch(1, nx, ny, seed=1234, device="cuda")
hF = torch.empty((S + 1, ny, nx), dtype=torch.float32)
hF[0::2] = I0[0].cpu(); hF[1::2] = I1[0].cpu()
h8 = torch.clamp(torch.round(hF), 0, 255).to(torch.uint8).pin_memory()
hW = h8.float().pin_memory()
u1 = torch.empty((S, ny, nx), dtype=torch.float32).pin_memory(); u2 = torch.empty_like(u1).pin_memory()
g = pkg.TVL1(0, max_batch=16, profiling=False); g.set_lanes(host_lanes=4)
def t(fn, reps=4):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return S / best
for rnd in range(3):
    a = t(lambda: g.solve_sequence_host_ptr(hW.data_ptr(), u1.data_ptr(), u2.data_ptr(), S + 1, nx, ny))
    b = t(lambda: g.solve_sequence_host_ptr(h8.data_ptr(), u1.data_ptr(), u2.data_ptr(), S + 1, nx, ny, dtype="uint8"))
    print("round %d: fp32 frames %.1f pairs/s, u8 frames %.1f pairs/s" % (rnd, a, b), flush=True)
