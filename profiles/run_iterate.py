"""Kernel-only driver for ncu: the fused iteration kernel on 32 resident 1080p pairs
(state + constants = 32 x 40 B x 2.07 Mpx = 2.65 GB, far larger than L2)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optical_flow_1_b200 as pkg

npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
ny = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 20
g = pkg.TVL1(0)
ms = g.bench_iterate(npairs, nx, ny, launches)
px = npairs * nx * ny * launches
print("k_iterate: %d pairs %dx%d, %d launches: %.3f ms/launch, %.1f GB/s algorithmic (64 B/px-iter)"
      % (npairs, nx, ny, launches, ms / launches, 64 * px / (ms / 1e3) / 1e9))
