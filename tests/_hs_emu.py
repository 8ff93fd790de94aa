"""Builds and binds tests/c/hs_schedule_emu.cpp (TEST INFRASTRUCTURE): the CPU replay of the
Horn-Schunck SOR kernel's schedule and the sequential fp32 sweep it is compared with."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

import _cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "hs_schedule_emu.cpp")
_lib = None


def load():
    global _lib
    if _lib is None:
        d = tempfile.mkdtemp(prefix="hs_emu_")
        so = os.path.join(d, "libhs_emu.so")
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-Wall", "-Wextra", "-Werror",
                        "-shared", "-o", so, SRC, "-lm"], check=True)
        lib = C.CDLL(so)
        vp, i, f, d_ = C.c_void_p, C.c_int, C.c_float, C.c_double
        lib.hs_emu_seq_sor.argtypes = [vp, vp, vp, vp, vp, i, i, f, d_, i, C.POINTER(d_)]
        lib.hs_emu_seq_sor.restype = i
        lib.hs_emu_wave_sor.argtypes = [vp, vp, vp, vp, vp, i, i, f, d_, i, i, i, i, i, i, C.c_uint, C.POINTER(d_)]
        lib.hs_emu_wave_sor.restype = i
        lib.hs_emu_pipe_wave_sor.argtypes = [vp, vp, vp, vp, vp, i, i, f, d_, i, i, i, i, i, i, i, C.c_uint,
                                             C.POINTER(d_), C.POINTER(i)]
        lib.hs_emu_pipe_wave_sor.restype = i
        lib.hs_emu_pairs_wave_sor.argtypes = lib.hs_emu_pipe_wave_sor.argtypes
        lib.hs_emu_pairs_wave_sor.restype = i
        _lib = lib
    return _lib


def system(nx, ny, seed):
    """Stored planes of the kernel -- I2wx, I2wy, rho_c (= -dif, src/horn_schunck_pyramidal.cpp:129-130)
    -- and a start flow, all fp32; plus the fp64 inputs they were made from."""
    x = _cases.hs_sor_inputs(nx, ny, seed)
    rho = -(x["I1"] - x["I2w"] + x["I2wx"] * x["u"] + x["I2wy"] * x["v"])
    f = lambda a: np.ascontiguousarray(a, np.float32)
    return f(x["I2wx"]), f(x["I2wy"]), f(rho), f(x["u"]), f(x["v"]), x


def run_seq(ix, iy, rho, u, v, alpha, tol, maxiter):
    """Sequential lexicographic sweep(s) in the kernel's fp32 arithmetic -> (u, v, sweeps, error)."""
    emu = load()
    u, v = u.copy(), v.copy()
    ny, nx = u.shape
    err = C.c_double()
    n = emu.hs_emu_seq_sor(u.ctypes.data, v.ctypes.data, ix.ctypes.data, iy.ctypes.data, rho.ctypes.data,
                           nx, ny, alpha * alpha, tol, maxiter, C.byref(err))
    return u, v, n, err.value


def run_wave(ix, iy, rho, u, v, alpha, tol, maxiter, P, nthreads, order, phase, land, seed=1):
    """The kernel's wavefront schedule replayed on the CPU under the given adversary."""
    emu = load()
    u, v = u.copy(), v.copy()
    ny, nx = u.shape
    err = C.c_double()
    n = emu.hs_emu_wave_sor(u.ctypes.data, v.ctypes.data, ix.ctypes.data, iy.ctypes.data, rho.ctypes.data,
                            nx, ny, alpha * alpha, tol, maxiter, P, nthreads, order, phase, land, seed,
                            C.byref(err))
    return u, v, n, err.value


def run_pipe_wave(ix, iy, rho, u, v, alpha, tol, maxiter, K, P, nthreads, order, phase, land, seed=1, pairs=False):
    """The PIPELINED schedule of k_hs_sor_pipe (hs_sor_pipe.h), or with pairs=True the two-columns-per-step
    schedule of hs_sor_pairs.h, replayed on the CPU -> (u, v, sweeps, error, sweeps replayed after a restore)."""
    emu = load()
    u, v = u.copy(), v.copy()
    ny, nx = u.shape
    err, rep = C.c_double(), C.c_int()
    fn = emu.hs_emu_pairs_wave_sor if pairs else emu.hs_emu_pipe_wave_sor
    n = fn(u.ctypes.data, v.ctypes.data, ix.ctypes.data, iy.ctypes.data, rho.ctypes.data,
                                 nx, ny, alpha * alpha, tol, maxiter, K, P, nthreads, order, phase, land, seed,
                                 C.byref(err), C.byref(rep))
    return u, v, n, err.value, rep.value
