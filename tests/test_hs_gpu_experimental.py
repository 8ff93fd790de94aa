"""GPU tests of kernels that have NOT run on a GPU yet (opt-in: HS_EXPERIMENTAL=1 pytest
tests/test_hs_gpu_experimental.py).  They are deliberately not marked `gpu`: the round-end GPU suite must
only contain tests that have passed on a B200.

k_hs_sor_pairs -- pipelined sweeps with two columns per thread-step (csrc/hs_sor_pairs.h).  Its step
functions are verified bit for bit by the CPU replay (tests/test_hs_schedule.py); the CUDA wrapper and the
pair-layout transposes are what these tests are for (hook code prefetch = -4 forces the kernel)."""
import os

import numpy as np
import pytest

import _hs_emu

pytestmark = pytest.mark.skipif(os.environ.get("HS_EXPERIMENTAL") != "1",
                                reason="experimental kernels: set HS_EXPERIMENTAL=1 on a GPU box")


@pytest.fixture(scope="module")
def gpu():
    import optical_flow_1_b200 as pkg
    g = pkg.HornSchunck(device=0)
    yield g
    g.close()


@pytest.mark.parametrize("nx,ny,sweeps", [(37, 29, 7), (64, 48, 7), (131, 70, 5), (33, 200, 5), (40, 1100, 3), (32, 3, 9),
                                          (1920, 1080, 3)])
def test_pairs_kernel_is_the_sequential_sweep_bitwise(gpu, nx, ny, sweeps):
    ix, iy, rho, u, v, _ = _hs_emu.system(nx, ny, seed=nx * 100 + ny)
    ru, rv, rn, rerr = _hs_emu.run_seq(ix, iy, rho, u, v, 7.0, 0.0, sweeps)
    gu, gv, gn, gerr = gpu.sor(ix, iy, rho, u, v, alpha=7.0, tol=0.0, maxiter=sweeps, prefetch=-4)
    assert gn == rn == sweeps
    assert np.array_equal(gu, ru) and np.array_equal(gv, rv), (np.abs(gu - ru).max(), np.abs(gv - rv).max())
    assert abs(gerr - rerr) <= 1e-9 * max(1.0, rerr)


@pytest.mark.parametrize("nx,ny", [(96, 80), (64, 300)])
@pytest.mark.parametrize("tol", [1e-1, 1e-2, 1e-3])
def test_pairs_kernel_stops_exactly(gpu, nx, ny, tol):
    ix, iy, rho, u, v, _ = _hs_emu.system(nx, ny, seed=3)
    ru, rv, rn, rerr = _hs_emu.run_seq(ix, iy, rho, u, v, 7.0, tol, 150)
    gu, gv, gn, gerr = gpu.sor(ix, iy, rho, u, v, alpha=7.0, tol=tol, maxiter=150, prefetch=-4)
    assert 1 < rn < 150 and gn == rn
    assert np.array_equal(gu, ru) and np.array_equal(gv, rv)
    assert abs(gerr - rerr) <= 1e-9 * max(1.0, rerr)
