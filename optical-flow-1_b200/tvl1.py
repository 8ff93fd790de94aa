"""Python mirror of the reference's TV-L1 interface on top of the C ABI (include/tvl1_b200.h).

The compute path is libtvl1_b200.so (hand-written sm_100a kernels).  There is no fallback: if
the library is missing or no CUDA device is present, construction raises.

Names follow the reference: `Dual_TVL1_optic_flow_multiscale` / `Dual_TVL1_optic_flow`
(src/tvl1flow.h:36-70), `zoom_size` / `zoom_out` / `zoom_in` (src/zoom.h), `gaussian`
(src/operators.h:128), `image_normalization_2` (src/utils.h:27).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

__all__ = ["TVL1", "TVL1Error", "Params", "Stats", "library_path", "build_library",
           "PAR_DEFAULTS", "clamp_nscales", "plan_chunks"]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libtvl1_b200.so")

# tvl1flow_main.cpp:24-33 (the CLI's nscales default is 100, clamped by image size at :185-188)
PAR_DEFAULTS = dict(tau=0.25, lam=0.15, theta=0.3, nscales=5, zfactor=0.5, warps=5, eps=0.01)

ERR_NAMES = {1: "TVL1_ERR_CUDA", 2: "TVL1_ERR_SIGMA", 3: "TVL1_ERR_ARG", 4: "TVL1_ERR_NODEVICE"}


class TVL1Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (ERR_NAMES.get(code, code), msg))
        self.code = code


class Params(C.Structure):
    _fields_ = [("tau", C.c_double), ("lam", C.c_double), ("theta", C.c_double),
                ("nscales", C.c_int), ("zfactor", C.c_double), ("warps", C.c_int),
                ("epsilon", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_ulonglong), ("iterate_launches", C.c_ulonglong),
                ("pixel_iterations", C.c_ulonglong), ("pixel_warps", C.c_ulonglong),
                ("iterate_ms", C.c_double), ("warp_ms", C.c_double), ("total_ms", C.c_double),
                ("pyramid_ms", C.c_double), ("zoom_in_ms", C.c_double), ("export_ms", C.c_double),
                ("host_syncs", C.c_ulonglong),
                ("level_pixel_iterations", C.c_ulonglong * 16),
                ("level_iterate_launches", C.c_ulonglong * 16),
                ("level_iterate_ms", C.c_double * 16),
                ("level_first_block_ms", C.c_double * 16)]

    def as_dict(self):
        out = {}
        for k, _ in self._fields_:
            v = getattr(self, k)
            out[k] = list(v) if hasattr(v, "__len__") else v
        return out


def library_path():
    return _LIB


def build_library(verbose=False):
    """Compile csrc/ for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)
    return _LIB


def clamp_nscales(nx, ny, nscales, zfactor):
    """The CLI's rule (tvl1flow_main.cpp:185-188): coarsest level not below about 16x16."""
    import math
    N = 1 + math.log(math.hypot(nx, ny) / 16.0) / math.log(1 / zfactor)
    return int(N) if N < nscales else nscales


def plan_chunks(npairs, max_batch):
    """Chunk sizes of a pinned host-buffer batch (include/tvl1_b200.h: tvl1_plan_chunks); needs no GPU."""
    lib = _load()
    lib.tvl1_plan_chunks.restype = C.c_int
    lib.tvl1_plan_chunks.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]
    n = lib.tvl1_plan_chunks(npairs, max_batch, None, 0)
    buf = (C.c_int * max(n, 1))()
    lib.tvl1_plan_chunks(npairs, max_batch, buf, n)
    return [int(buf[k]) for k in range(n)]


def _load():
    if not os.path.exists(_LIB):
        raise FileNotFoundError(
            "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % _LIB)
    lib = C.CDLL(_LIB)
    lib.tvl1_last_error.restype = C.c_char_p
    lib.tvl1_last_error.argtypes = [C.c_void_p]
    return lib


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


class TVL1:
    """One solver context = one GPU + one stream (one per thread / per rank)."""

    def __init__(self, device=0, max_batch=None, profiling=False):
        self.lib = _load()
        self.ctx = C.c_void_p()
        rc = self.lib.tvl1_create(C.c_int(device), C.byref(self.ctx))
        if rc:
            raise TVL1Error(rc, self.lib.tvl1_last_error(None).decode())
        self.device = device
        if max_batch:
            self.lib.tvl1_set_max_batch(self.ctx, C.c_int(int(max_batch)))
        if profiling:
            self.set_profiling(True)

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.tvl1_destroy(self.ctx)
            self.ctx = None

    __del__ = close

    def _ck(self, rc):
        if rc:
            raise TVL1Error(rc, self.lib.tvl1_last_error(self.ctx).decode())

    def set_profiling(self, on):
        self._ck(self.lib.tvl1_set_profiling(self.ctx, C.c_int(1 if on else 0)))

    def set_lanes(self, host_lanes=3, dev_lanes=2):
        self._ck(self.lib.tvl1_set_lanes(self.ctx, C.c_int(int(host_lanes)), C.c_int(int(dev_lanes))))

    def set_max_batch(self, pairs):
        self._ck(self.lib.tvl1_set_max_batch(self.ctx, C.c_int(int(pairs))))

    def stream(self):
        """cudaStream_t (as int) that all work of this context is issued on."""
        self.lib.tvl1_get_stream.restype = C.c_void_p
        return int(self.lib.tvl1_get_stream(self.ctx) or 0)

    def blocked_levels(self):
        """Bit mask of the pyramid levels whose big launches run two iterations per launch (tvl1_get_blocked_levels)."""
        self.lib.tvl1_get_blocked_levels.restype = C.c_uint
        return int(self.lib.tvl1_get_blocked_levels(self.ctx))

    def stats(self):
        s = Stats()
        self._ck(self.lib.tvl1_get_stats(self.ctx, C.byref(s)))
        return s.as_dict()

    @staticmethod
    def _params(tau, lam, theta, nscales, zfactor, warps, eps):
        return Params(tau, lam, theta, int(nscales), zfactor, int(warps), eps)

    # -- solver ----------------------------------------------------------------------------------
    def Dual_TVL1_optic_flow_multiscale(self, I0, I1, tau=0.25, lam=0.15, theta=0.3, nscales=5,
                                        zfactor=0.5, warps=5, eps=0.01, verbose=False):
        """src/tvl1flow.cpp:219-328.  I0, I1: (ny, nx) or (npairs, ny, nx), float32 or float64
        host arrays.  Returns (u1, u2, iters, errs); iters/errs are (..., nscales, warps) with the
        COARSEST level first (the order of the reference's verbose output)."""
        I0 = np.asarray(I0)
        dt = np.float64 if I0.dtype == np.float64 else np.float32
        I0 = np.ascontiguousarray(I0, dt)
        I1 = np.ascontiguousarray(I1, dt)
        assert I0.shape == I1.shape and I0.ndim in (2, 3)
        batched = I0.ndim == 3
        npairs = I0.shape[0] if batched else 1
        ny, nx = I0.shape[-2:]
        u1 = np.empty(I0.shape, dt)
        u2 = np.empty(I0.shape, dt)
        iters = np.zeros((npairs, nscales, warps), np.int32)
        errs = np.zeros((npairs, nscales, warps), np.float64)
        prm = self._params(tau, lam, theta, nscales, zfactor, warps, eps)
        fn = self.lib.tvl1_solve_batch_f64 if dt == np.float64 else self.lib.tvl1_solve_batch_f32
        self._ck(fn(self.ctx, C.c_int(npairs), _fp(I0), _fp(I1), _fp(u1), _fp(u2), C.c_int(nx),
                    C.c_int(ny), C.byref(prm), _fp(iters), _fp(errs)))
        if verbose:
            sizes = [(nx, ny)]
            for _ in range(1, nscales):
                sizes.append(self.zoom_size(sizes[-1][0], sizes[-1][1], zfactor))
            for b in range(npairs):
                for s in range(nscales - 1, -1, -1):
                    sys.stderr.write("Scale %d: %dx%d\n" % (s, sizes[s][0], sizes[s][1]))
                    for w in range(warps):
                        sys.stderr.write("Warping: %d, Iterations: %d, Error: %f\n"
                                         % (w, iters[b, nscales - 1 - s, w], errs[b, nscales - 1 - s, w]))
        if not batched:
            iters, errs = iters[0], errs[0]
        return u1, u2, iters, errs

    def Dual_TVL1_optic_flow(self, I0, I1, u1, u2, tau=0.25, lam=0.15, theta=0.3, warps=5, eps=0.01):
        """src/tvl1flow.cpp:46-212: one level; (u1, u2) is the initial flow.  Returns
        (u1, u2, iters[warps], errs[warps])."""
        I0 = np.asarray(I0)
        dt = np.float64 if I0.dtype == np.float64 else np.float32
        I0 = np.ascontiguousarray(I0, dt)
        I1 = np.ascontiguousarray(I1, dt)
        u1 = np.array(u1, dtype=dt, order="C", copy=True)
        u2 = np.array(u2, dtype=dt, order="C", copy=True)
        ny, nx = I0.shape
        iters = np.zeros(warps, np.int32)
        errs = np.zeros(warps, np.float64)
        prm = self._params(tau, lam, theta, 1, 0.5, warps, eps)
        fn = self.lib.tvl1_single_scale_f64 if dt == np.float64 else self.lib.tvl1_single_scale_f32
        self._ck(fn(self.ctx, _fp(I0), _fp(I1), _fp(u1), _fp(u2), C.c_int(nx), C.c_int(ny),
                    C.byref(prm), _fp(iters), _fp(errs)))
        return u1, u2, iters, errs

    def solve_batch_device(self, dI0, dI1, du1, du2, npairs, nx, ny, tau=0.25, lam=0.15, theta=0.3,
                           nscales=5, zfactor=0.5, warps=5, eps=0.01, want_iters=False):
        """Device-resident batch: dI0.. are integer device addresses of dense float32
        [npairs][ny][nx] buffers (e.g. torch.Tensor.data_ptr())."""
        prm = self._params(tau, lam, theta, nscales, zfactor, warps, eps)
        iters = errs = None
        ip = ep = None
        if want_iters:
            iters = np.zeros((npairs, nscales, warps), np.int32)
            errs = np.zeros((npairs, nscales, warps), np.float64)
            ip, ep = _fp(iters), _fp(errs)
        self._ck(self.lib.tvl1_solve_batch_dev_f32(
            self.ctx, C.c_int(npairs), C.c_void_p(dI0), C.c_void_p(dI1), C.c_void_p(du1),
            C.c_void_p(du2), C.c_int(nx), C.c_int(ny), C.byref(prm), ip, ep))
        return iters, errs

    def solve_batch_host_ptr(self, pI0, pI1, pu1, pu2, npairs, nx, ny, dtype=np.float32, iters=None, errs=None, **kw):
        """Host-resident batch by raw address (e.g. pinned torch tensors): the drop-in call with
        H2D / D2H inside.  `iters` / `errs`: optional int32 / float64 arrays [npairs, nscales, warps]."""
        p = dict(PAR_DEFAULTS)
        p.update(kw)
        prm = self._params(p["tau"], p["lam"], p["theta"], p["nscales"], p["zfactor"], p["warps"], p["eps"])
        fn = self.lib.tvl1_solve_batch_f64 if np.dtype(dtype) == np.float64 else self.lib.tvl1_solve_batch_f32
        self._ck(fn(self.ctx, C.c_int(npairs), C.c_void_p(pI0), C.c_void_p(pI1), C.c_void_p(pu1),
                    C.c_void_p(pu2), C.c_int(nx), C.c_int(ny), C.byref(prm),
                    None if iters is None else iters.ctypes.data_as(C.c_void_p),
                    None if errs is None else errs.ctypes.data_as(C.c_void_p)))

    def plan_chunks(self, npairs, max_batch):
        """Chunk sizes a pinned host-buffer batch is cut into (tvl1_plan_chunks)."""
        return plan_chunks(npairs, max_batch)

    def solve_sequence(self, frames, tau=0.25, lam=0.15, theta=0.3, nscales=5, zfactor=0.5, warps=5,
                       eps=0.01):
        """A video: frames (F, ny, nx), float32 or float64 -> flows of the F-1 consecutive pairs,
        (u1, u2, iters[F-1, nscales, warps], errs) -- the reference CLI's call
        (src/tvl1flow_main.cpp:203-206) looped over consecutive frames, each frame uploaded once."""
        frames = np.asarray(frames)
        if frames.dtype == np.uint8:                      # 8-bit video: bytes over PCIe, fp32 flows
            frames = np.ascontiguousarray(frames)
            F, ny, nx = frames.shape
            u1 = np.empty((max(F - 1, 0), ny, nx), np.float32)
            u2 = np.empty_like(u1)
            iters = np.zeros((max(F - 1, 0), nscales, warps), np.int32)
            errs = np.zeros((max(F - 1, 0), nscales, warps), np.float64)
            prm = self._params(tau, lam, theta, nscales, zfactor, warps, eps)
            p = lambda a: a.ctypes.data_as(C.c_void_p)
            self._ck(self.lib.tvl1_solve_sequence_u8(self.ctx, C.c_int(F), p(frames), p(u1), p(u2), C.c_int(nx), C.c_int(ny),
                                                     C.byref(prm), p(iters), p(errs)))
            return u1, u2, iters, errs
        dt = np.float64 if frames.dtype == np.float64 else np.float32
        frames = np.ascontiguousarray(frames, dt)
        F, ny, nx = frames.shape
        u1 = np.empty((max(F - 1, 0), ny, nx), dt)
        u2 = np.empty_like(u1)
        iters = np.zeros((max(F - 1, 0), nscales, warps), np.int32)
        errs = np.zeros((max(F - 1, 0), nscales, warps), np.float64)
        prm = self._params(tau, lam, theta, nscales, zfactor, warps, eps)
        fn = self.lib.tvl1_solve_sequence_f64 if dt == np.float64 else self.lib.tvl1_solve_sequence_f32
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self._ck(fn(self.ctx, C.c_int(F), p(frames), p(u1), p(u2), C.c_int(nx), C.c_int(ny), C.byref(prm),
                    iters.ctypes.data_as(C.POINTER(C.c_int)), errs.ctypes.data_as(C.POINTER(C.c_double))))
        return u1, u2, iters, errs

    def solve_sequence_host_ptr(self, pframes, pu1, pu2, nframes, nx, ny, dtype=np.float32, iters=None, errs=None, **kw):
        """Frame sequence by raw host address (pinned buffers)."""
        p = dict(PAR_DEFAULTS)
        p.update(kw)
        prm = self._params(p["tau"], p["lam"], p["theta"], p["nscales"], p["zfactor"], p["warps"], p["eps"])
        fn = (self.lib.tvl1_solve_sequence_u8 if np.dtype(dtype) == np.uint8 else
              self.lib.tvl1_solve_sequence_f64 if np.dtype(dtype) == np.float64 else self.lib.tvl1_solve_sequence_f32)
        self._ck(fn(self.ctx, C.c_int(nframes), C.c_void_p(pframes), C.c_void_p(pu1), C.c_void_p(pu2),
                    C.c_int(nx), C.c_int(ny), C.byref(prm),
                    None if iters is None else iters.ctypes.data_as(C.c_void_p),
                    None if errs is None else errs.ctypes.data_as(C.c_void_p)))

    # -- row-band mode: one image pair over several GPUs -------------------------------------------
    @staticmethod
    def _prefer_torch_nccl():
        """The library binds NCCL at run time and takes the copy the process has already loaded.  In a
        Python process that copy should be torch's own (torch.distributed is the plumbing around the band
        mode): import torch first, so that a later `import torch` does not find an older system libnccl
        under the same soname."""
        try:
            import torch  # noqa: F401
        except ImportError:
            pass

    def band_unique_id(self):
        self._prefer_torch_nccl()
        buf = (C.c_ubyte * 128)()
        rc = self.lib.tvl1_band_unique_id(buf)
        if rc:
            raise TVL1Error(rc, self.lib.tvl1_last_error(None).decode())
        return bytes(buf)

    def band_init(self, rank, world, unique_id):
        self._prefer_torch_nccl()
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        self._ck(self.lib.tvl1_band_init(self.ctx, C.c_int(rank), C.c_int(world), buf))

    def band_set_exchange(self, mode):
        """'peer' (default: NVLink stores + mailboxes inside the iteration kernel) or 'nccl'."""
        self._ck(self.lib.tvl1_band_set_exchange(self.ctx, C.c_int(1 if mode == "nccl" else 0)))

    def band_exchange_mode(self):
        return {1: "peer", 0: "nccl", -1: None}[self.lib.tvl1_band_exchange_mode(self.ctx)]

    def band_rows(self, ny, rank, world):
        a, b = C.c_int(), C.c_int()
        self.lib.tvl1_band_rows(C.c_int(ny), C.c_int(rank), C.c_int(world), C.byref(a), C.byref(b))
        return a.value, b.value

    def band_solve(self, I0, I1, min_split_rows=512, tau=0.25, lam=0.15, theta=0.3, nscales=5,
                   zfactor=0.5, warps=5, eps=0.01):
        """Collective over the ranks of band_init: same full (ny, nx) float32 images on every rank;
        returns the full flow on every rank: (u1, u2, iters[nscales, warps], errs)."""
        I0, I1 = self._f32(I0), self._f32(I1)
        ny, nx = I0.shape
        u1, u2 = np.empty_like(I0), np.empty_like(I0)
        iters = np.zeros((nscales, warps), np.int32)
        errs = np.zeros((nscales, warps), np.float64)
        prm = self._params(tau, lam, theta, nscales, zfactor, warps, eps)
        self._ck(self.lib.tvl1_band_solve_f32(self.ctx, _fp(I0), _fp(I1), _fp(u1), _fp(u2), C.c_int(nx),
                                              C.c_int(ny), C.byref(prm), C.c_int(min_split_rows),
                                              _fp(iters), _fp(errs)))
        return u1, u2, iters, errs

    def band_solve_host_ptr(self, pI0, pI1, pu1, pu2, nx, ny, min_split_rows=512, **kw):
        """Banded solve by raw HOST addresses (e.g. pinned torch tensors); collective."""
        p = dict(PAR_DEFAULTS)
        p.update(kw)
        prm = self._params(p["tau"], p["lam"], p["theta"], p["nscales"], p["zfactor"], p["warps"], p["eps"])
        self._ck(self.lib.tvl1_band_solve_f32(self.ctx, C.c_void_p(pI0), C.c_void_p(pI1), C.c_void_p(pu1),
                                              C.c_void_p(pu2), C.c_int(nx), C.c_int(ny), C.byref(prm),
                                              C.c_int(min_split_rows), None, None))

    def band_solve_device(self, dI0, dI1, du1, du2, nx, ny, min_split_rows=512, tau=0.25, lam=0.15,
                          theta=0.3, nscales=5, zfactor=0.5, warps=5, eps=0.01):
        prm = self._params(tau, lam, theta, nscales, zfactor, warps, eps)
        iters = np.zeros((nscales, warps), np.int32)
        errs = np.zeros((nscales, warps), np.float64)
        self._ck(self.lib.tvl1_band_solve_dev_f32(self.ctx, C.c_void_p(dI0), C.c_void_p(dI1), C.c_void_p(du1),
                                                  C.c_void_p(du2), C.c_int(nx), C.c_int(ny), C.byref(prm),
                                                  C.c_int(min_split_rows), _fp(iters), _fp(errs)))
        return iters, errs

    # -- per-kernel hooks ------------------------------------------------------------------------
    def zoom_size(self, nx, ny, factor):
        a, b = C.c_int(), C.c_int()
        self.lib.tvl1_zoom_size(C.c_int(nx), C.c_int(ny), C.byref(a), C.byref(b), C.c_double(factor))
        return a.value, b.value

    @staticmethod
    def _f32(a):
        return np.ascontiguousarray(a, np.float32)

    def image_normalization_2(self, I0, I1):
        I0, I1 = self._f32(I0), self._f32(I1)
        ny, nx = I0.shape
        o0, o1 = np.empty_like(I0), np.empty_like(I1)
        self._ck(self.lib.tvl1_normalize_f32(self.ctx, _fp(I0), _fp(I1), _fp(o0), _fp(o1),
                                             C.c_int(nx), C.c_int(ny)))
        return o0, o1

    def gaussian(self, I, sigma):
        I = self._f32(I)
        ny, nx = I.shape
        out = np.empty_like(I)
        self._ck(self.lib.tvl1_gaussian_f32(self.ctx, _fp(I), _fp(out), C.c_int(nx), C.c_int(ny),
                                            C.c_double(sigma)))
        return out

    def zoom_out(self, I, factor):
        I = self._f32(I)
        ny, nx = I.shape
        nxx, nyy = self.zoom_size(nx, ny, factor)
        out = np.empty((nyy, nxx), np.float32)
        self._ck(self.lib.tvl1_zoom_out_f32(self.ctx, _fp(I), _fp(out), C.c_int(nx), C.c_int(ny),
                                            C.c_double(factor)))
        return out

    def zoom_in(self, I, nxx, nyy, scale=1.0):
        I = self._f32(I)
        ny, nx = I.shape
        out = np.empty((nyy, nxx), np.float32)
        self._ck(self.lib.tvl1_zoom_in_f32(self.ctx, _fp(I), _fp(out), C.c_int(nx), C.c_int(ny),
                                           C.c_int(nxx), C.c_int(nyy), C.c_double(scale)))
        return out

    def warp_precompute(self, I0, I1, u1, u2):
        """src/tvl1flow.cpp:84,94-109 fused -> dict(I1wx, I1wy, rho_c, grad)"""
        I0, I1, u1, u2 = (self._f32(a) for a in (I0, I1, u1, u2))
        ny, nx = I0.shape
        out = [np.empty_like(I0) for _ in range(4)]
        self._ck(self.lib.tvl1_warp_f32(self.ctx, _fp(I0), _fp(I1), _fp(u1), _fp(u2), C.c_int(nx),
                                        C.c_int(ny), *[_fp(o) for o in out]))
        return dict(I1wx=out[0], I1wy=out[1], rho_c=out[2], grad=out[3])

    def iterate(self, u1, u2, p11, p12, p21, p22, rho_c, I1wx, I1wy, grad, tau, lam, theta, iters):
        st = [np.array(a, dtype=np.float32, order="C", copy=True) for a in (u1, u2, p11, p12, p21, p22)]
        cs = [self._f32(a) for a in (rho_c, I1wx, I1wy, grad)]
        ny, nx = st[0].shape
        errs = np.zeros(max(iters, 1), np.float64)
        self._ck(self.lib.tvl1_iterate_f32(self.ctx, *[_fp(a) for a in st], *[_fp(a) for a in cs],
                                           C.c_int(nx), C.c_int(ny), C.c_double(tau), C.c_double(lam),
                                           C.c_double(theta), C.c_int(iters), _fp(errs)))
        return (*st, errs[:iters])

    def iterate_resident(self, u1, u2, p11, p12, p21, p22, rho_c, I1wx, I1wy, tau, lam, theta, eps,
                         max_iter, cluster=0):
        """Cluster-resident kernel: whole while loop on chip.  Returns
        (u1,u2,p11,p12,p21,p22, n, errs[n], cluster_size_used)."""
        st = [np.array(a, dtype=np.float32, order="C", copy=True) for a in (u1, u2, p11, p12, p21, p22)]
        cs = [self._f32(a) for a in (rho_c, I1wx, I1wy)]
        ny, nx = st[0].shape
        errs = np.zeros(max_iter, np.float64)
        n, c = C.c_int(), C.c_int()
        self._ck(self.lib.tvl1_iterate_resident_f32(
            self.ctx, *[_fp(a) for a in st], *[_fp(a) for a in cs], C.c_int(nx), C.c_int(ny),
            C.c_double(tau), C.c_double(lam), C.c_double(theta), C.c_double(eps), C.c_int(max_iter),
            C.c_int(cluster), C.byref(n), _fp(errs), C.byref(c)))
        return (*st, n.value, errs[:n.value], c.value)

    def iterate_loop(self, u1, u2, p11, p12, p21, p22, rho_c, I1wx, I1wy, tau, lam, theta, eps, max_iter,
                     temporal_blocking=1):
        """One warp step's while loop through the streaming kernels.  Returns
        (u1,u2,p11,p12,p21,p22, n, last_error, launches)."""
        st = [np.array(a, dtype=np.float32, order="C", copy=True) for a in (u1, u2, p11, p12, p21, p22)]
        cs = [self._f32(a) for a in (rho_c, I1wx, I1wy)]
        ny, nx = st[0].shape
        n, launches, err = C.c_int(), C.c_int(), C.c_double()
        self._ck(self.lib.tvl1_iterate_loop_f32(
            self.ctx, *[_fp(a) for a in st], *[_fp(a) for a in cs], C.c_int(nx), C.c_int(ny),
            C.c_double(tau), C.c_double(lam), C.c_double(theta), C.c_double(eps), C.c_int(max_iter),
            C.c_int(temporal_blocking), C.byref(n), C.byref(err), C.byref(launches)))
        return (*st, n.value, err.value, launches.value)

    def bench_iterate(self, npairs, nx, ny, launches):
        ms = C.c_double()
        self._ck(self.lib.tvl1_bench_iterate(self.ctx, C.c_int(npairs), C.c_int(nx), C.c_int(ny),
                                             C.c_int(launches), C.byref(ms)))
        return ms.value
