"""Kernel-only rates of the iteration kernels on resident 1080p pairs (tvl1_bench_iterate): one iteration per launch
(k_iterate_t1), two per launch in registers (k_iterate_t2, TVL1_BENCH_TB=2), up to four in shared memory (k_iterate_tb,
TVL1_BENCH_TB=1).  A launch advances every pair by 1 / 2 / 4 iterations; GB/s is algorithmic (64 B per pixel-iteration)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optical_flow_1_b200 as pkg
launches = 20
for npairs, nx, ny in [(32, 1920, 1080), (128, 1920, 1080), (128, 960, 540)]:
    for mode, per in ((0, 1), (2, 2), (1, 4)):
        os.environ["TVL1_BENCH_TB"] = str(mode)
        g = pkg.TVL1(0)
        ms = g.bench_iterate(npairs, nx, ny, launches)
        g.close()
        px = npairs * nx * ny * launches * per
        print("mode %d (%d iteration(s) per launch): %3d pairs %dx%d: %.3f ms/launch, %.3f ms/iteration, %.1f GB/s algorithmic"
              % (mode, per, npairs, nx, ny, ms / launches, ms / launches / per, 64 * px / (ms / 1e3) / 1e9), flush=True)
