"""Design prototypes of Horn-Schunck SOR schedules, modelled on plain arrays on the
CPU and compared bit for bit with the sequential sweep, sweep count included:

* tests/c/hs_pipeline_emu.cpp -- PIPELINED sweeps: sweep n of row i at time n*L + 2i + j with L = nx + 2, all
  rows busy all the time, snapshot + replay for the exact stopping rule.  Built since: k_hs_sor_pipe.
* tests/c/hs_pairs_emu.cpp -- the next step, not yet a kernel: TWO columns per thread-step on top of it
  (pair c = columns 2c, 2c+1 at time n*L + 2i + c)."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

import _hs_emu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


SOURCES = {"hs_emu_pipe_sor": "hs_pipeline_emu.cpp", "hs_emu_pairs_sor": "hs_pairs_emu.cpp"}


@pytest.fixture(scope="module", params=sorted(SOURCES))
def pipe(request):
    """The model under test: (library, entry point)."""
    name = request.param
    so = os.path.join(tempfile.mkdtemp(prefix="hs_pipe_"), "lib%s.so" % name)
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-Wall", "-Wextra", "-Werror", "-shared",
                    "-o", so, os.path.join(ROOT, "tests", "c", SOURCES[name]), "-lm"], check=True)
    lib = C.CDLL(so)
    vp, i, f, d = C.c_void_p, C.c_int, C.c_float, C.c_double
    fn = getattr(lib, name)
    fn.argtypes = [vp, vp, vp, vp, vp, i, i, f, d, i, i, i, i, C.c_uint, C.POINTER(d), C.POINTER(i), C.POINTER(i)]
    fn.restype = i
    return fn


def run_pipe(fn, ix, iy, rho, u, v, alpha, tol, maxiter, K, nthreads, order, seed=1):
    u, v = u.copy(), v.copy()
    ny, nx = u.shape
    err, rep, spec = C.c_double(), C.c_int(), C.c_int()
    n = fn(u.ctypes.data, v.ctypes.data, ix.ctypes.data, iy.ctypes.data, rho.ctypes.data, nx, ny, alpha * alpha, tol,
           maxiter, K, nthreads, order, seed, C.byref(err), C.byref(rep), C.byref(spec))
    return u, v, n, err.value, rep.value, spec.value


SHAPES = [(3, 3), (4, 3), (3, 4), (5, 3), (3, 9), (9, 3), (4, 4), (5, 7), (8, 5), (12, 9), (37, 29), (64, 48), (23, 70),
          (70, 23)]


@pytest.mark.parametrize("nx,ny", SHAPES)
def test_pipelined_sweeps_equal_the_sequential_loop_fixed_count(pipe, nx, ny):
    """tol = 0: exactly maxiter sweeps (the maxiter stop needs no restore)."""
    ix, iy, rho, u, v, _ = _hs_emu.system(nx, ny, seed=nx * 100 + ny)
    for maxiter in (1, 2, 5, 11):
        ref = _hs_emu.run_seq(ix, iy, rho, u, v, 7.0, 0.0, maxiter)
        for nthreads in sorted({1, 3, ny, ny + 5}):
            for order in range(3):
                got = run_pipe(pipe, ix, iy, rho, u, v, 7.0, 0.0, maxiter, 4, nthreads, order, seed=order + maxiter)
                key = (maxiter, nthreads, order)
                assert got[2] == ref[2] == maxiter, key
                assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), key
                assert abs(got[3] - ref[3]) <= 1e-12 * max(1.0, ref[3]), key
                assert got[4] == 0, key


@pytest.mark.parametrize("nx,ny", [(12, 9), (37, 29), (64, 48), (23, 70), (70, 23)])
@pytest.mark.parametrize("K", [1, 3, 4, 8])
def test_pipelined_sweeps_stop_exactly_where_the_sequential_loop_stops(pipe, nx, ny, K):
    """Stops by TOL: rows had run ahead, the state is restored from the snapshot and the rest replayed."""
    ix, iy, rho, u, v, _ = _hs_emu.system(nx, ny, seed=7 * nx + ny)
    seen = set()
    for tol in (3e-1, 1e-1, 3e-2, 1e-2, 3e-3, 1e-3):
        ref = _hs_emu.run_seq(ix, iy, rho, u, v, 7.0, tol, 150)
        if ref[2] in seen:
            continue
        seen.add(ref[2])
        for order in range(3):
            got = run_pipe(pipe, ix, iy, rho, u, v, 7.0, tol, 150, K, max(1, ny // 2), order, seed=order)
            assert got[2] == ref[2], (tol, order, got[2], ref[2])
            assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), (tol, order)
            assert abs(got[3] - ref[3]) <= 1e-12 * max(1.0, ref[3])
            if ref[2] < 150:
                assert got[5] >= 1                     # speculative sweeps had been started ...
                assert got[4] < max(K, (2 * ny + nx) // 8 + 2)     # ... and fewer than a snapshot period were replayed
    assert len(seen) >= 3
