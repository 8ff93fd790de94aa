"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the CPU oracle.

Two interchangeable back ends with the same Python surface (class ``CpuTvl1``):

* ``port``      -- oracle/tvl1_oracle.c, the C restatement (always available; built on demand)
* ``reference`` -- oracle/_ref/libof_ref_*.so, the unmodified reference TUs compiled by
                   ``make -C oracle ref`` (present when /root/reference was available at build
                   time; the built .so travels to the GPU box, the sources do not)

Only tests/, ``__graft_entry__.smoke()`` and the cpu_baseline / ``--impl reference`` legs of
bench.py may import this module.  Nothing under optical-flow-1_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = "/root/reference"

_c_int_p = C.POINTER(C.c_int)
_c_double_p = C.POINTER(C.c_double)


def build(ref=True, quiet=True):
    """Compile the oracle port and, when the reference tree is present, oracle/_ref."""
    targets = ["all"]
    if ref and os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        targets.append("ref")
    subprocess.run(["make", "-C", HERE, "-j4"] + targets, check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _path(kind, dtype):
    suffix = "f64" if np.dtype(dtype) == np.float64 else "f32"
    if kind == "port":
        return os.path.join(HERE, "libtvl1_oracle_%s.so" % suffix)
    return os.path.join(HERE, "_ref", "libof_ref_%s.so" % suffix)


def available(kind, dtype=np.float64):
    return os.path.exists(_path(kind, dtype))


class CpuTvl1:
    """CPU TV-L1 (oracle port or compiled reference) behind one numpy interface."""

    def __init__(self, kind="port", dtype=np.float64):
        assert kind in ("port", "reference")
        self.kind = kind
        self.dtype = np.dtype(dtype)
        path = _path(kind, dtype)
        if not os.path.exists(path):
            build(ref=(kind == "reference"))
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.pfx = "orc_" if kind == "port" else "ref_"
        assert self._f("sizeof_pix")() == self.dtype.itemsize
        self._f("max_threads").restype = C.c_int

    # -- helpers -------------------------------------------------------------------------------
    def _f(self, name):
        return getattr(self.lib, self.pfx + name)

    def _arr(self, a):
        return np.ascontiguousarray(a, dtype=self.dtype)

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    def max_threads(self):
        return int(self._f("max_threads")())

    def set_threads(self, n):
        self._f("set_threads")(C.c_int(int(n)))

    # -- (a) pyramid ---------------------------------------------------------------------------
    def normalize(self, I0, I1):
        I0, I1 = self._arr(I0), self._arr(I1)
        o0, o1 = np.empty_like(I0), np.empty_like(I1)
        self._f("normalize")(self._p(I0), self._p(I1), self._p(o0), self._p(o1), C.c_int(I0.size))
        return o0, o1

    def gaussian(self, I, sigma):
        I = self._arr(I).copy()
        ny, nx = I.shape
        rc = self._f("gaussian")(self._p(I), C.c_int(nx), C.c_int(ny), C.c_double(sigma))
        if rc:
            raise RuntimeError("GaussianSmooth: sigma too large")
        return I

    def zoom_size(self, nx, ny, factor):
        a, b = C.c_int(), C.c_int()
        self._f("zoom_size")(C.c_int(nx), C.c_int(ny), C.byref(a), C.byref(b), C.c_double(factor))
        return a.value, b.value

    def zoom_out(self, I, factor):
        I = self._arr(I)
        ny, nx = I.shape
        nxx, nyy = self.zoom_size(nx, ny, factor)
        out = np.empty((nyy, nxx), self.dtype)
        rc = self._f("zoom_out")(self._p(I), self._p(out), C.c_int(nx), C.c_int(ny), C.c_double(factor))
        if rc:
            raise RuntimeError("GaussianSmooth: sigma too large")
        return out

    def zoom_in(self, I, nxx, nyy):
        I = self._arr(I)
        ny, nx = I.shape
        out = np.empty((nyy, nxx), self.dtype)
        self._f("zoom_in")(self._p(I), self._p(out), C.c_int(nx), C.c_int(ny), C.c_int(nxx), C.c_int(nyy))
        return out

    # -- (b) warp ------------------------------------------------------------------------------
    def centered_gradient(self, I):
        I = self._arr(I)
        ny, nx = I.shape
        dx, dy = np.empty_like(I), np.empty_like(I)
        self._f("centered_gradient")(self._p(I), self._p(dx), self._p(dy), C.c_int(nx), C.c_int(ny))
        return dx, dy

    def warp(self, I, u, v, border_out=True):
        I, u, v = self._arr(I), self._arr(u), self._arr(v)
        ny, nx = I.shape
        out = np.empty_like(I)
        self._f("warp")(self._p(I), self._p(u), self._p(v), self._p(out), C.c_int(nx), C.c_int(ny),
                        C.c_int(1 if border_out else 0))
        return out

    def warp_precompute(self, I0, I1, u1, u2):
        """src/tvl1flow.cpp:84,94-109 -> dict(I1w, I1wx, I1wy, rho_c, grad)"""
        I1x, I1y = self.centered_gradient(I1)
        I1w = self.warp(I1, u1, u2)
        I1wx = self.warp(I1x, u1, u2)
        I1wy = self.warp(I1y, u1, u2)
        t = self.dtype.type
        u1, u2, I0 = self._arr(u1), self._arr(u2), self._arr(I0)
        if self.dtype == np.float64:
            grad = I1wx * I1wx + I1wy * I1wy
            rho_c = I1w - I1wx * u1 - I1wy * u2 - I0
        else:
            # the float build squares in float, adds in double, stores float (tvl1flow.cpp:100-104)
            grad = ((I1wx * I1wx).astype(np.float64) + (I1wy * I1wy).astype(np.float64)).astype(t)
            rho_c = I1w - I1wx * u1 - I1wy * u2 - I0
        return dict(I1w=I1w, I1wx=I1wx, I1wy=I1wy, rho_c=rho_c, grad=grad)

    # -- (c) iteration -------------------------------------------------------------------------
    def divergence(self, v1, v2):
        v1, v2 = self._arr(v1), self._arr(v2)
        ny, nx = v1.shape
        out = np.empty_like(v1)
        self._f("divergence")(self._p(v1), self._p(v2), self._p(out), C.c_int(nx), C.c_int(ny))
        return out

    def forward_gradient(self, f):
        f = self._arr(f)
        ny, nx = f.shape
        fx, fy = np.empty_like(f), np.empty_like(f)
        self._f("forward_gradient")(self._p(f), self._p(fx), self._p(fy), C.c_int(nx), C.c_int(ny))
        return fx, fy

    def iterate(self, u1, u2, p11, p12, p21, p22, rho_c, I1wx, I1wy, grad, tau, lam, theta, iters):
        """Exactly `iters` passes of src/tvl1flow.cpp:114-181 (port only). Returns
        (u1,u2,p11,p12,p21,p22, errs[iters])."""
        assert self.kind == "port"
        st = [self._arr(a).copy() for a in (u1, u2, p11, p12, p21, p22)]
        cs = [self._arr(a) for a in (rho_c, I1wx, I1wy, grad)]
        ny, nx = st[0].shape
        errs = np.zeros(max(iters, 1), np.float64)
        self._f("iterate")(*[self._p(a) for a in st], *[self._p(a) for a in cs], C.c_int(nx),
                           C.c_int(ny), C.c_double(tau), C.c_double(lam), C.c_double(theta),
                           C.c_int(iters), errs.ctypes.data_as(_c_double_p))
        return (*st, errs[:iters])

    def single_scale(self, I0, I1, u1, u2, tau=0.25, lam=0.15, theta=0.3, warps=5, eps=0.01):
        I0, I1 = self._arr(I0), self._arr(I1)
        u1, u2 = self._arr(u1).copy(), self._arr(u2).copy()
        ny, nx = I0.shape
        iters = np.zeros(warps, np.int32)
        errs = np.zeros(warps, np.float64)
        args = [self._p(I0), self._p(I1), self._p(u1), self._p(u2), C.c_int(nx), C.c_int(ny),
                C.c_double(tau), C.c_double(lam), C.c_double(theta), C.c_int(warps), C.c_double(eps),
                iters.ctypes.data_as(_c_int_p), errs.ctypes.data_as(_c_double_p)]
        if self.kind == "port":
            self._f("single_scale")(*args)
        else:
            n = self._f("single_scale_iters")(*args, C.c_int(warps))
            assert n == warps, n
        return u1, u2, iters, errs

    def multiscale(self, I0, I1, tau=0.25, lam=0.15, theta=0.3, nscales=5, zfactor=0.5, warps=5,
                   eps=0.01, want_iters=True):
        """Returns (u1, u2, iters[nscales, warps], errs[nscales, warps]); row 0 is the COARSEST
        scale (the order of the reference's verbose output)."""
        I0, I1 = self._arr(I0), self._arr(I1)
        ny, nx = I0.shape
        u1 = np.empty((ny, nx), self.dtype)
        u2 = np.empty((ny, nx), self.dtype)
        iters = np.zeros(nscales * warps, np.int32)
        errs = np.zeros(nscales * warps, np.float64)
        base = [self._p(I0), self._p(I1), self._p(u1), self._p(u2), C.c_int(nx), C.c_int(ny),
                C.c_double(tau), C.c_double(lam), C.c_double(theta), C.c_int(nscales),
                C.c_double(zfactor), C.c_int(warps), C.c_double(eps)]
        if self.kind == "port":
            rc = self._f("multiscale")(*base, iters.ctypes.data_as(_c_int_p),
                                       errs.ctypes.data_as(_c_double_p))
            if rc:
                raise RuntimeError("GaussianSmooth: sigma too large")
        elif want_iters:
            n = self._f("multiscale_iters")(*base, iters.ctypes.data_as(_c_int_p),
                                            errs.ctypes.data_as(_c_double_p), C.c_int(iters.size))
            assert n == iters.size, n
        else:
            self._f("multiscale")(*base, C.c_int(0))
        return u1, u2, iters.reshape(nscales, warps), errs.reshape(nscales, warps)

    # -- (d) pyramidal Horn-Schunck, src/horn_schunck_pyramidal.cpp ----------------------------------
    # The reference's SOR sweep updates u, v in place under an OpenMP parallel-for (:148-158): only
    # its one-thread result is well defined.  The reference back end is therefore run with one
    # thread (and the previous thread count restored); the port is sequential by construction.
    def hs_system(self, I1, I2w, I2wx, I2wy, u, v, alpha):
        """src/horn_schunck_pyramidal.cpp:127-137 (port only) -> (Au, Av, Du, Dv, D)."""
        assert self.kind == "port"
        arrs = [self._arr(a) for a in (I1, I2w, I2wx, I2wy, u, v)]
        outs = [np.empty_like(arrs[0]) for _ in range(5)]
        self._f("hs_system")(*[self._p(a) for a in arrs], *[self._p(a) for a in outs],
                             C.c_double(alpha * alpha), C.c_int(arrs[0].size))
        return tuple(outs)

    def hs_sor(self, Au, Av, Du, Dv, D, u, v, alpha, tol, maxiter):
        """The SOR loop of one warp step, :139-231 (port only) -> (u, v, niter, error)."""
        assert self.kind == "port"
        cs = [self._arr(a) for a in (Au, Av, Du, Dv, D)]
        u, v = self._arr(u).copy(), self._arr(v).copy()
        ny, nx = u.shape
        err = C.c_double()
        f = self._f("hs_sor")
        f.restype = C.c_int
        n = f(*[self._p(a) for a in cs], self._p(u), self._p(v), C.c_double(alpha * alpha), C.c_int(nx),
              C.c_int(ny), C.c_double(tol), C.c_int(maxiter), C.byref(err))
        return u, v, int(n), err.value

    def hs_single_scale(self, I1, I2, u, v, alpha=7.0, warps=10, tol=1e-4, maxiter=150):
        """horn_schunck_optical_flow (u, v in/out) -> (u, v, iters[warps], errs[warps])."""
        I1, I2 = self._arr(I1), self._arr(I2)
        u, v = self._arr(u).copy(), self._arr(v).copy()
        ny, nx = I1.shape
        iters = np.zeros(warps, np.int32)
        errs = np.zeros(warps, np.float64)
        args = [self._p(I1), self._p(I2), self._p(u), self._p(v), C.c_int(nx), C.c_int(ny),
                C.c_double(alpha), C.c_int(warps), C.c_double(tol), C.c_int(maxiter),
                iters.ctypes.data_as(_c_int_p), errs.ctypes.data_as(_c_double_p)]
        if self.kind == "port":
            self._f("hs_single_scale")(*args)
        else:
            old = self.max_threads()
            self.set_threads(1)
            try:
                n = self._f("hs_single_scale_iters")(*args, C.c_int(warps))
            finally:
                self.set_threads(old)
            assert n == warps, n
        return u, v, iters, errs

    def hs_multiscale(self, I1, I2, alpha=7.0, nscales=5, zfactor=0.5, warps=10, tol=1e-4, maxiter=150,
                      threads=1):
        """horn_schunck_pyramidal -> (u, v, iters[nscales, warps], errs[nscales, warps]); row 0 is the
        coarsest scale.  `threads` > 1 (reference back end only) runs the reference's racy parallel
        sweep: used by the CPU baseline timing, never for parity."""
        I1, I2 = self._arr(I1), self._arr(I2)
        ny, nx = I1.shape
        u = np.empty((ny, nx), self.dtype)
        v = np.empty((ny, nx), self.dtype)
        iters = np.zeros(nscales * warps, np.int32)
        errs = np.zeros(nscales * warps, np.float64)
        args = [self._p(I1), self._p(I2), self._p(u), self._p(v), C.c_int(nx), C.c_int(ny),
                C.c_double(alpha), C.c_int(nscales), C.c_double(zfactor), C.c_int(warps), C.c_double(tol),
                C.c_int(maxiter), iters.ctypes.data_as(_c_int_p), errs.ctypes.data_as(_c_double_p)]
        if self.kind == "port":
            rc = self._f("hs_multiscale")(*args)
            if rc:
                raise RuntimeError("GaussianSmooth: sigma too large")
        else:
            old = self.max_threads()
            self.set_threads(threads)
            try:
                n = self._f("hs_multiscale_iters")(*args, C.c_int(iters.size))
            finally:
                self.set_threads(old)
            assert n == iters.size, n
        return u, v, iters.reshape(nscales, warps), errs.reshape(nscales, warps)


UPSTREAM_C99 = os.path.join(HERE, "_ref", "libtvl1flow3_c99.so")


def upstream_c99_available():
    return os.path.exists(UPSTREAM_C99)


def c99_signature(lib):
    """Declares the C-linkage prototypes of 3rdparty/tvl1flow_3/tvl1flow_lib.c:45-59 and :299-314 on
    a loaded library (the upstream build, or the CUDA library that exports the same names)."""
    vp, f, i, b = C.c_void_p, C.c_float, C.c_int, C.c_bool
    lib.Dual_TVL1_optic_flow_multiscale.argtypes = [vp, vp, vp, vp, i, i, f, f, f, i, f, i, f, b]
    lib.Dual_TVL1_optic_flow_multiscale.restype = None
    lib.Dual_TVL1_optic_flow.argtypes = [vp, vp, vp, vp, i, i, f, f, f, i, f, b]
    lib.Dual_TVL1_optic_flow.restype = None
    return lib


def c99_multiscale(lib, I0, I1, tau=0.25, lam=0.15, theta=0.3, nscales=5, zfactor=0.5, warps=5, eps=0.01):
    """Calls lib's C-linkage Dual_TVL1_optic_flow_multiscale(float ...) and returns (u1, u2)."""
    I0 = np.ascontiguousarray(I0, np.float32).copy()
    I1 = np.ascontiguousarray(I1, np.float32).copy()
    ny, nx = I0.shape
    u = np.zeros((2, ny, nx), np.float32)
    lib.Dual_TVL1_optic_flow_multiscale(I0.ctypes.data, I1.ctypes.data, u[0].ctypes.data, u[1].ctypes.data,
                                        nx, ny, tau, lam, theta, nscales, zfactor, warps, eps, False)
    return u[0], u[1]


# ---- TV-L1 with occlusions -----------------------------------------------------------
OCC_DEFAULTS = dict(lam=0.15, alpha=0.01, beta=0.15, theta=0.3, nscales=5, zfactor=0.5, warps=2, eps=0.01)


def occ_available(kind, dtype=np.float64):
    suffix = "f64" if np.dtype(dtype) == np.float64 else "f32"
    if kind == "port":
        return os.path.exists(os.path.join(HERE, "libtvl1_oracle_%s.so" % suffix))
    return os.path.exists(os.path.join(HERE, "_ref", "libocc_ref_%s.so" % suffix))


class CpuOcc:
    """src/tvl1occflow*.cpp on the CPU: the C restatement (``port``, oracle/tvl1_oracle.c section (e)) or the
    unmodified reference objects behind oracle/occ_ref_shim.cpp (``reference``; zero-filling new[], see
    that file).  Images are (ny, nx) arrays; I_1 = frame before I0, filtI0 = the image g is taken from."""

    def __init__(self, kind="port", dtype=np.float64):
        assert kind in ("port", "reference")
        self.kind, self.dtype = kind, np.dtype(dtype)
        suffix = "f64" if self.dtype == np.float64 else "f32"
        path = (os.path.join(HERE, "libtvl1_oracle_%s.so" % suffix) if kind == "port"
                else os.path.join(HERE, "_ref", "libocc_ref_%s.so" % suffix))
        if not os.path.exists(path):
            build(ref=(kind == "reference"))
        self.lib = C.CDLL(path)
        self.pfx = "orc_occ_" if kind == "port" else "occ_ref_"
        if kind == "reference":
            assert self.lib.occ_ref_sizeof_pix() == self.dtype.itemsize
            assert self.lib.occ_ref_new_is_zeroing() == 1
        else:
            assert self.lib.orc_sizeof_pix() == self.dtype.itemsize

    def _arr(self, a):
        return np.ascontiguousarray(a, dtype=self.dtype)

    _p = staticmethod(lambda a: a.ctypes.data_as(C.c_void_p))

    def set_threads(self, n):
        (self.lib.orc_set_threads if self.kind == "port" else self.lib.occ_ref_set_threads)(C.c_int(int(n)))

    def multiscale(self, I_1, I0, I1, filtI0=None, lam=0.15, alpha=0.01, beta=0.15, theta=0.3, nscales=5,
                   zfactor=0.5, warps=2, eps=0.01):
        """-> (u1, u2, chi, iters[nscales, warps] coarsest level first, errs)."""
        I_1, I0, I1 = self._arr(I_1), self._arr(I0), self._arr(I1)
        f = self._arr(I0 if filtI0 is None else filtI0)
        ny, nx = I0.shape
        u1, u2, chi = (np.zeros_like(I0) for _ in range(3))
        it = np.zeros((nscales, warps), np.int32)
        er = np.zeros((nscales, warps), np.float64)
        rc = getattr(self.lib, self.pfx + "multiscale")(
            self._p(I_1), self._p(I0), self._p(I1), self._p(f), self._p(u1), self._p(u2), self._p(chi),
            C.c_int(nx), C.c_int(ny), C.c_double(lam), C.c_double(alpha), C.c_double(beta), C.c_double(theta),
            C.c_int(nscales), C.c_double(zfactor), C.c_int(warps), C.c_double(eps), self._p(it), self._p(er))
        if self.kind == "port" and rc:
            raise RuntimeError("GaussianSmooth: sigma too large")
        return u1, u2, chi, it, er

    def single_scale(self, I_1, I0, I1, filtI0, u1, u2, chi, lam=0.15, alpha=0.01, beta=0.15, theta=0.3, warps=2,
                     eps=0.01):
        """Dual_TVL1_optic_flow of src/tvl1occflow.cpp:144-330 (restatement only): one level from the given
        flow and occlusion map -> (u1, u2, chi, iters[warps], errs[warps])."""
        assert self.kind == "port"
        I_1, I0, I1 = self._arr(I_1), self._arr(I0), self._arr(I1)
        f = self._arr(I0 if filtI0 is None else filtI0)
        u1, u2, chi = self._arr(u1).copy(), self._arr(u2).copy(), self._arr(chi).copy()
        ny, nx = I0.shape
        it = np.zeros(warps, np.int32)
        er = np.zeros(warps, np.float64)
        self.lib.orc_occ_single_scale(self._p(I_1), self._p(I0), self._p(I1), self._p(f), self._p(u1), self._p(u2),
                                      self._p(chi), C.c_int(nx), C.c_int(ny), C.c_double(lam), C.c_double(alpha),
                                      C.c_double(beta), C.c_double(theta), C.c_int(warps), C.c_double(eps),
                                      self._p(it), self._p(er))
        return u1, u2, chi, it, er

    def rof_box(self, u, f, p1, p2, g, lam, omega=1.25, niter=10):
        """Scalar_ROF_BoxCellCentered -> (u, p1, p2) after niter sweeps."""
        u, p1, p2 = self._arr(u).copy(), self._arr(p1).copy(), self._arr(p2).copy()
        f, g = self._arr(f), self._arr(g)
        ny, nx = u.shape
        getattr(self.lib, self.pfx + "rof_box")(self._p(u), self._p(f), self._p(p1), self._p(p2), self._p(g),
                                                C.c_double(lam), C.c_double(omega), C.c_int(nx), C.c_int(ny),
                                                C.c_int(niter))
        return u, p1, p2

    def median3(self, a):
        a = self._arr(a).copy()
        ny, nx = a.shape
        (self.lib.orc_median3 if self.kind == "port" else self.lib.occ_ref_median3)(self._p(a), C.c_int(nx), C.c_int(ny))
        return a
