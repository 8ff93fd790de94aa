"""Python mirror of the reference's pyramidal Horn-Schunck interface on top of the C ABI
(include/hs_b200.h).  Names follow the reference: `horn_schunck_pyramidal` /
`horn_schunck_optical_flow` (src/horn_schunck.h:15-48).  Same library, same context type and the
same rule as the TV-L1 mirror: no fallback -- without the CUDA library or a device, construction
raises."""
import ctypes as C
import sys

import numpy as np

from .tvl1 import TVL1, _fp

__all__ = ["HornSchunck", "HsParams", "HS_DEFAULTS", "hs_clamp_nscales"]

# src/horn_schunck_pyramidal_main.cpp:25-30
HS_DEFAULTS = dict(alpha=7.0, nscales=10, zfactor=0.5, warps=10, tol=1e-4, maxiter=150)


class HsParams(C.Structure):
    _fields_ = [("alpha", C.c_double), ("nscales", C.c_int), ("zfactor", C.c_double),
                ("warps", C.c_int), ("tol", C.c_double), ("maxiter", C.c_int)]


def hs_clamp_nscales(nx, ny, nscales, zfactor):
    """The CLI's rule (src/horn_schunck_pyramidal_main.cpp:136-143): no level below about 16x16."""
    import math
    N = 1 + math.log(math.hypot(nx, ny) / 16.0) / math.log(1 / zfactor)
    return int(N) if N < nscales else nscales


class HornSchunck(TVL1):
    """One solver context (one GPU + one stream); also offers everything `TVL1` does."""

    @staticmethod
    def _hs_params(alpha, nscales, zfactor, warps, tol, maxiter):
        return HsParams(alpha, int(nscales), zfactor, int(warps), tol, int(maxiter))

    def horn_schunck_pyramidal(self, I1, I2, alpha=7.0, nscales=10, zfactor=0.5, warps=10, tol=1e-4,
                               maxiter=150, verbose=False):
        """src/horn_schunck_pyramidal.cpp:258-370.  I1, I2: (ny, nx) or (npairs, ny, nx), float32 or
        float64 host arrays.  Returns (u, v, iters, errs); iters/errs are (..., nscales, warps) with
        the COARSEST level first (the order of the reference's verbose output)."""
        I1 = np.asarray(I1)
        dt = np.float64 if I1.dtype == np.float64 else np.float32
        I1 = np.ascontiguousarray(I1, dt)
        I2 = np.ascontiguousarray(I2, dt)
        assert I1.shape == I2.shape and I1.ndim in (2, 3)
        batched = I1.ndim == 3
        npairs = I1.shape[0] if batched else 1
        ny, nx = I1.shape[-2:]
        u = np.empty(I1.shape, dt)
        v = np.empty(I1.shape, dt)
        iters = np.zeros((npairs, nscales, warps), np.int32)
        errs = np.zeros((npairs, nscales, warps), np.float64)
        prm = self._hs_params(alpha, nscales, zfactor, warps, tol, maxiter)
        if dt == np.float64:
            assert not batched, "batches go through the float32 entry point"
            self._ck(self.lib.hs_solve_f64(self.ctx, _fp(I1), _fp(I2), _fp(u), _fp(v), C.c_int(nx), C.c_int(ny),
                                           C.byref(prm), _fp(iters), _fp(errs)))
        else:
            self._ck(self.lib.hs_solve_batch_f32(self.ctx, C.c_int(npairs), _fp(I1), _fp(I2), _fp(u), _fp(v),
                                                 C.c_int(nx), C.c_int(ny), C.byref(prm), _fp(iters), _fp(errs)))
        if verbose:
            sizes = [(nx, ny)]
            for _ in range(1, nscales):
                sizes.append(self.zoom_size(sizes[-1][0], sizes[-1][1], zfactor))
            for b in range(npairs):
                for s in range(nscales - 1, -1, -1):
                    sys.stderr.write("Scale: %d %dx%d\n" % (s, sizes[s][0], sizes[s][1]))
                    for w in range(warps):
                        sys.stderr.write("Warping %d:Iterations %d (%g)\n"
                                         % (w, iters[b, nscales - 1 - s, w], errs[b, nscales - 1 - s, w]))
        if not batched:
            iters, errs = iters[0], errs[0]
        return u, v, iters, errs

    def horn_schunck_optical_flow(self, I1, I2, u, v, alpha=7.0, warps=10, tol=1e-4, maxiter=150):
        """src/horn_schunck_pyramidal.cpp:78-249: one level; (u, v) is the initial flow.  Returns
        (u, v, iters[warps], errs[warps])."""
        I1 = np.asarray(I1)
        dt = np.float64 if I1.dtype == np.float64 else np.float32
        I1 = np.ascontiguousarray(I1, dt)
        I2 = np.ascontiguousarray(I2, dt)
        u = np.array(u, dtype=dt, order="C", copy=True)
        v = np.array(v, dtype=dt, order="C", copy=True)
        ny, nx = I1.shape
        iters = np.zeros(warps, np.int32)
        errs = np.zeros(warps, np.float64)
        prm = self._hs_params(alpha, 1, 0.5, warps, tol, maxiter)
        fn = self.lib.hs_single_scale_f64 if dt == np.float64 else self.lib.hs_single_scale_f32
        self._ck(fn(self.ctx, _fp(I1), _fp(I2), _fp(u), _fp(v), C.c_int(nx), C.c_int(ny), C.byref(prm),
                    _fp(iters), _fp(errs)))
        return u, v, iters, errs

    def hs_solve_batch_device(self, dI1, dI2, du, dv, npairs, nx, ny, want_iters=False, **kw):
        """Device-resident batch: integer device addresses of dense float32 [npairs][ny][nx] buffers."""
        p = dict(HS_DEFAULTS)
        p.update(kw)
        prm = self._hs_params(p["alpha"], p["nscales"], p["zfactor"], p["warps"], p["tol"], p["maxiter"])
        iters = errs = None
        ip = ep = None
        if want_iters:
            iters = np.zeros((npairs, p["nscales"], p["warps"]), np.int32)
            errs = np.zeros((npairs, p["nscales"], p["warps"]), np.float64)
            ip, ep = _fp(iters), _fp(errs)
        self._ck(self.lib.hs_solve_batch_dev_f32(self.ctx, C.c_int(npairs), C.c_void_p(dI1), C.c_void_p(dI2),
                                                 C.c_void_p(du), C.c_void_p(dv), C.c_int(nx), C.c_int(ny),
                                                 C.byref(prm), ip, ep))
        return iters, errs

    def sor(self, I2wx, I2wy, rho_c, u, v, alpha=7.0, tol=1e-4, maxiter=150, prefetch=-1):
        """Test hook: the SOR loop of one warp step (src/horn_schunck_pyramidal.cpp:139-231) on a given
        system.  Returns (u, v, sweeps, error)."""
        f = lambda a: np.ascontiguousarray(a, np.float32)
        I2wx, I2wy, rho_c = f(I2wx), f(I2wy), f(rho_c)
        u = np.array(u, dtype=np.float32, order="C", copy=True)
        v = np.array(v, dtype=np.float32, order="C", copy=True)
        ny, nx = u.shape
        n, e = C.c_int(), C.c_double()
        self._ck(self.lib.hs_sor_f32(self.ctx, _fp(I2wx), _fp(I2wy), _fp(rho_c), _fp(u), _fp(v), C.c_int(nx),
                                     C.c_int(ny), C.c_double(alpha), C.c_double(tol), C.c_int(maxiter),
                                     C.c_int(prefetch), C.byref(n), C.byref(e)))
        return u, v, n.value, e.value
