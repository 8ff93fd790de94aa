"""Python mirror of the reference's TV-L1 + occlusions interface on top of the C ABI (include/occ_b200.h).
Names follow the reference: the seven-plane `Dual_TVL1_optic_flow_multiscale` / `Dual_TVL1_optic_flow`
of src/tvl1occflow.h:63-129.  fp64 in, fp64 out (the reference's shipped pixel type); no fallback --
without the CUDA library or a device, construction raises."""
import ctypes as C
import sys

import numpy as np

from .tvl1 import _fp, _load

__all__ = ["TVL1Occ", "OccError", "OccParams", "OccStats", "OCC_DEFAULTS", "occ_clamp_nscales"]

# src/tvl1occflow_constants.h:14-23 (the CLI's nscales default is 100, clamped by image size)
OCC_DEFAULTS = dict(lam=0.15, alpha=0.01, beta=0.15, theta=0.3, nscales=5, zfactor=0.5, warps=2, eps=0.01)

ERR_NAMES = {1: "OCC_ERR_CUDA", 2: "OCC_ERR_SIGMA", 3: "OCC_ERR_ARG", 4: "OCC_ERR_NODEVICE"}


class OccError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (ERR_NAMES.get(code, code), msg))
        self.code = code


class OccParams(C.Structure):
    _fields_ = [("lam", C.c_double), ("alpha", C.c_double), ("beta", C.c_double), ("theta", C.c_double),
                ("nscales", C.c_int), ("zfactor", C.c_double), ("warps", C.c_int), ("epsilon", C.c_double)]


class OccStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_ulonglong), ("outer_iterations", C.c_ulonglong),
                ("box_sweeps", C.c_ulonglong), ("box_cell_updates", C.c_ulonglong),
                ("chi_pixel_iterations", C.c_ulonglong), ("host_syncs", C.c_ulonglong),
                ("ms_total", C.c_double), ("ms_pyramid", C.c_double), ("ms_warp", C.c_double),
                ("ms_box", C.c_double), ("ms_chi", C.c_double), ("ms_other", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def occ_clamp_nscales(nx, ny, nscales, zfactor):
    """The CLI's rule (src/tvl1occflow_main.cpp:191-196)."""
    lib = _load()
    lib.occ_clamp_nscales.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double]
    return lib.occ_clamp_nscales(nx, ny, nscales, zfactor)


class TVL1Occ:
    """One solver context = one GPU + one stream."""

    def __init__(self, device=0, profiling=False, max_batch=None):
        self.lib = _load()
        self.lib.occ_last_error.restype = C.c_char_p
        self.lib.occ_last_error.argtypes = [C.c_void_p]
        self.lib.occ_get_stream.restype = C.c_void_p
        self.lib.occ_get_stream.argtypes = [C.c_void_p]
        self.ctx = C.c_void_p()
        rc = self.lib.occ_create(C.c_int(device), C.byref(self.ctx))
        if rc:
            raise OccError(rc, self.lib.occ_last_error(None).decode())
        self.device = device
        if profiling:
            self.lib.occ_set_profiling(self.ctx, C.c_int(1))
        if max_batch:
            self.lib.occ_set_max_batch(self.ctx, C.c_int(int(max_batch)))

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.occ_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise OccError(rc, self.lib.occ_last_error(self.ctx).decode())

    def stream(self):
        return self.lib.occ_get_stream(self.ctx)

    def stats(self):
        st = OccStats()
        self._ck(self.lib.occ_get_stats(self.ctx, C.byref(st)))
        return st.as_dict()

    @staticmethod
    def _params(lam, alpha, beta, theta, nscales, zfactor, warps, eps):
        return OccParams(lam, alpha, beta, theta, int(nscales), zfactor, int(warps), eps)

    def Dual_TVL1_optic_flow_multiscale(self, I_1, I0, I1, filtI0=None, lam=0.15, alpha=0.01, beta=0.15, theta=0.3,
                                        nscales=5, zfactor=0.5, warps=2, eps=0.01, verbose=False):
        """src/tvl1occflow.cpp:335-482.  Images: (ny, nx) or (ntriples, ny, nx) host arrays (converted to
        float64).  Returns (u1, u2, chi, iters, errs); chi is 0 / 1; iters / errs are (..., nscales, warps)
        with the COARSEST level first."""
        I0 = np.ascontiguousarray(I0, np.float64)
        I_1 = np.ascontiguousarray(I_1, np.float64)
        I1 = np.ascontiguousarray(I1, np.float64)
        f = None if filtI0 is None else np.ascontiguousarray(filtI0, np.float64)
        assert I0.shape == I_1.shape == I1.shape and I0.ndim in (2, 3)
        batched = I0.ndim == 3
        n = I0.shape[0] if batched else 1
        ny, nx = I0.shape[-2:]
        u1, u2, chi = (np.empty(I0.shape, np.float64) for _ in range(3))
        iters = np.zeros((n, nscales, warps), np.int32)
        errs = np.zeros((n, nscales, warps), np.float64)
        prm = self._params(lam, alpha, beta, theta, nscales, zfactor, warps, eps)
        self._ck(self.lib.occ_solve_batch_f64(self.ctx, C.c_int(n), _fp(I_1), _fp(I0), _fp(I1),
                                              None if f is None else _fp(f), _fp(u1), _fp(u2), _fp(chi),
                                              C.c_int(nx), C.c_int(ny), C.byref(prm), _fp(iters), _fp(errs)))
        if verbose:
            for b in range(n):
                for k in range(nscales):
                    for w in range(warps):
                        sys.stderr.write("Warping: %d, Iterations: %d, Error: %e\n" % (w, iters[b, k, w], errs[b, k, w]))
        if not batched:
            iters, errs = iters[0], errs[0]
        return u1, u2, chi, iters, errs

    def Dual_TVL1_optic_flow(self, I_1, I0, I1, filtI0, u1, u2, chi, lam=0.15, alpha=0.01, beta=0.15, theta=0.3,
                             warps=2, eps=0.01):
        """src/tvl1occflow.cpp:144-330: one level; (u1, u2, chi) are the initial values.  Returns
        (u1, u2, chi, iters[warps], errs[warps]); chi is not thresholded."""
        a = lambda x: np.ascontiguousarray(x, np.float64)
        I_1, I0, I1 = a(I_1), a(I0), a(I1)
        f = None if filtI0 is None else a(filtI0)
        u1, u2, chi = (np.array(x, dtype=np.float64, order="C", copy=True) for x in (u1, u2, chi))
        ny, nx = I0.shape
        iters = np.zeros(warps, np.int32)
        errs = np.zeros(warps, np.float64)
        prm = self._params(lam, alpha, beta, theta, 1, 0.5, warps, eps)
        self._ck(self.lib.occ_single_scale_f64(self.ctx, _fp(I_1), _fp(I0), _fp(I1), None if f is None else _fp(f),
                                               _fp(u1), _fp(u2), _fp(chi), C.c_int(nx), C.c_int(ny), C.byref(prm),
                                               _fp(iters), _fp(errs)))
        return u1, u2, chi, iters, errs

    def solve_batch_device(self, dI_1, dI0, dI1, dfilt, du1, du2, dchi, n, nx, ny, want_iters=False, **kw):
        """Device-resident batch: integer device addresses of dense float64 [n][ny][nx] buffers (dfilt may be 0)."""
        p = dict(OCC_DEFAULTS)
        p.update(kw)
        prm = self._params(p["lam"], p["alpha"], p["beta"], p["theta"], p["nscales"], p["zfactor"], p["warps"], p["eps"])
        iters = errs = None
        ip = ep = None
        if want_iters:
            iters = np.zeros((n, p["nscales"], p["warps"]), np.int32)
            errs = np.zeros((n, p["nscales"], p["warps"]), np.float64)
            ip, ep = _fp(iters), _fp(errs)
        self._ck(self.lib.occ_solve_batch_dev_f64(self.ctx, C.c_int(n), C.c_void_p(dI_1), C.c_void_p(dI0),
                                                  C.c_void_p(dI1), C.c_void_p(dfilt or None), C.c_void_p(du1),
                                                  C.c_void_p(du2), C.c_void_p(dchi), C.c_int(nx), C.c_int(ny),
                                                  C.byref(prm), ip, ep))
        return iters, errs

    def rof_box(self, u, f, p1, p2, g, lam, omega=1.25, niter=10):
        """Test hook: Scalar_ROF_BoxCellCentered (src/tvl1occflow_tv_rof_box.cpp:25-645) -> (u, p1, p2)."""
        u, p1, p2 = (np.array(x, dtype=np.float64, order="C", copy=True) for x in (u, p1, p2))
        f, g = np.ascontiguousarray(f, np.float64), np.ascontiguousarray(g, np.float64)
        ny, nx = u.shape
        self._ck(self.lib.occ_rof_box_f64(self.ctx, _fp(u), _fp(f), _fp(p1), _fp(p2), _fp(g), C.c_double(lam),
                                          C.c_double(omega), C.c_int(nx), C.c_int(ny), C.c_int(niter)))
        return u, p1, p2

    def median3(self, a):
        """Test hook: me_median_filtering, window 3 (src/utils.cpp:151-213)."""
        a = np.array(a, dtype=np.float64, order="C", copy=True)
        ny, nx = a.shape
        self._ck(self.lib.occ_median3_f64(self.ctx, _fp(a), C.c_int(nx), C.c_int(ny)))
        return a
