"""Single-pair latency through the host-buffer C ABI (the drop-in call): median of 10 solves."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optical_flow_1_b200 as pkg

g = pkg.TVL1(0)
for (nx, ny, kw) in [(640, 480, {}), (1024, 436, {}), (1920, 1080, {}),
                     (3840, 2160, dict(nscales=6, warps=10, eps=0.001))]:
    I0, I1 = pkg.synth.make_pair(nx, ny, seed=1234)
    for dt in (np.float32, np.float64):
        a, b = I0.astype(dt), I1.astype(dt)
        ts = []
        for rep in range(6 if nx > 2000 else 12):
            t = time.perf_counter()
            u1, u2, it, er = g.Dual_TVL1_optic_flow_multiscale(a, b, **kw)
            ts.append(time.perf_counter() - t)
        st = g.stats()
        print("%dx%d %s: median %.2f ms (min %.2f)  iterations/level %s  launches %d"
              % (nx, ny, dt.__name__, 1e3 * np.median(ts[2:]), 1e3 * min(ts), it.sum(axis=1).tolist(), st["kernel_launches"]))
