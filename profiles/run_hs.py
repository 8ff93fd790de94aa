"""Pyramidal Horn-Schunck driver: one device-resident batch of N synthetic pairs solved
twice (the first solve warms up), device times by kernel group and the algorithmic bandwidth of the
SOR kernel.  Also the driver for ncu captures of k_hs_sor.

    python profiles/run_hs.py [npairs nx ny nscales warps maxiter tol] [--json out.json]

Algorithmic bytes of the sweep: 28 B per pixel and sweep (read u, v, I2wx, I2wy, rho_c; write u, v --
the system of src/horn_schunck_pyramidal.cpp:127-137 is formed on the fly, the reference reads
Au, Av, Du, Dv, D, u, v and writes u, v: 72 B in fp64)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_1_b200 as pkg

argv = [a for a in sys.argv[1:] if not a.startswith("--")]
out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
if out_json:
    argv.remove(out_json)
npairs = int(argv[0]) if len(argv) > 0 else 32
nx = int(argv[1]) if len(argv) > 1 else 1920
ny = int(argv[2]) if len(argv) > 2 else 1080
kw = dict(pkg.HS_DEFAULTS)
if len(argv) > 5:
    kw.update(nscales=int(argv[3]), warps=int(argv[4]), maxiter=int(argv[5]))
if len(argv) > 6:
    kw.update(tol=float(argv[6]))
kw["nscales"] = pkg.hs_clamp_nscales(nx, ny, kw["nscales"], kw["zfactor"])
I1, I2 = pkg.synth.make_batch_torch(npairs, nx, ny, seed=1234, device="cuda")
u, v = torch.empty_like(I1), torch.empty_like(I1)
torch.cuda.synchronize()
g = pkg.HornSchunck(0, max_batch=npairs, profiling=True)
for rep in range(2):
    it, er = g.hs_solve_batch_device(I1.data_ptr(), I2.data_ptr(), u.data_ptr(), v.data_ptr(), npairs, nx, ny,
                                     want_iters=True, **kw)
    st = g.stats()
sweep_bytes = 28.0 * st["pixel_iterations"]
res = dict(prefetch_env=os.environ.get('HS_PREFETCH'), npairs=npairs, nx=nx, ny=ny, params=kw, total_ms=st["total_ms"], sor_ms=st["iterate_ms"],
           warp_ms=st["warp_ms"], pyramid_ms=st["pyramid_ms"], zoom_in_ms=st["zoom_in_ms"],
           kernel_launches=st["kernel_launches"], sor_launches=st["iterate_launches"],
           pixel_sweeps=st["pixel_iterations"], pairs_per_s=npairs / (st["total_ms"] * 1e-3),
           sor_algorithmic_GBps=sweep_bytes / (st["iterate_ms"] * 1e-3) / 1e9 if st["iterate_ms"] else None,
           level_sor_ms=[round(x, 3) for x in st["level_iterate_ms"][:kw["nscales"]]],
           level_pixel_sweeps=st["level_pixel_iterations"][:kw["nscales"]],
           sweeps_pair0=it[0].tolist(), finite=bool(torch.isfinite(u).all() and torch.isfinite(v).all()),
           flow_absmax=float(max(u.abs().max(), v.abs().max())))
print(json.dumps(res))
if out_json:
    with open(out_json, "w") as f:
        json.dump(res, f, indent=1)
