"""Temporally blocked iteration kernel alone: 40 iterations (10 blocks of 4) on one 3840x2160 state,
next to the same 40 iterations through the streaming kernel (one per launch)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optical_flow_1_b200 as pkg

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 3840
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 2160
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rs = np.random.RandomState(1)
f = lambda lo, hi: rs.uniform(lo, hi, (ny, nx)).astype(np.float32)
u1, u2 = f(-3, 3), f(-3, 3)
p = [f(-1, 1) for _ in range(4)]
ix, iy, rho = f(-20, 20), f(-20, 20), f(-30, 30)
g = pkg.TVL1(0)
for mode, name in ((2, "k_iterate_tb (blocks of 4)"), (0, "k_iterate_t1 (1 per launch)")):
    best = 1e9
    for rep in range(3):
        t = time.perf_counter()
        out = g.iterate_loop(u1, u2, *p, rho, ix, iy, 0.25, 0.15, 0.3, -1.0, iters, mode)
        best = min(best, time.perf_counter() - t)
    print("%s: %d iterations in %d launches (wall incl. host copies %.1f ms)" % (name, out[6], out[8], 1e3 * best))
