"""The wavefront schedule of the Horn-Schunck SOR kernel, replayed on the CPU.

tests/c/hs_schedule_emu.cpp compiles the product's own per-thread step functions
(optical-flow-1_b200/csrc/hs_sor_step.h -- the text nvcc compiles into k_hs_sor) with g++ and drives
them like the kernel does (one barrier per time step, threads owning interleaved rows, asynchronous
copies), while the test plays the adversary: thread order inside a step, fetch / update phase order,
and the moment an asynchronous copy reads and lands.  The schedule is correct iff the result equals
the sequential lexicographic sweep of src/horn_schunck_pyramidal.cpp:144-230 bit for bit in every case.
A second test ties that sequential fp32 sweep to the fp64 oracle.
"""
import numpy as np
import pytest

import _hs_emu
from _hs_emu import run_pipe_wave, run_seq, run_wave, system


SHAPES = [(3, 3), (4, 3), (3, 4), (5, 3), (3, 9), (9, 3), (4, 4), (5, 7), (8, 5), (12, 9), (37, 29), (64, 48), (23, 70)]


@pytest.mark.parametrize("nx,ny", SHAPES)
def test_wave_schedule_equals_sequential_sweep(nx, ny):
    ix, iy, rho, u, v, _ = system(nx, ny, seed=nx * 100 + ny)
    ref = run_seq(ix, iy, rho, u, v, 7.0, 0.0, 5)
    assert ref[2] == 5 and np.isfinite(ref[0]).all()
    for P in range(4):
        for nthreads in sorted({1, 2, 3, max(1, ny // 2), ny, ny + 5}):
            for order in range(3):
                for phase in range(3):
                    for land in range(3):
                        got = run_wave(ix, iy, rho, u, v, 7.0, 0.0, 5, P, nthreads, order, phase, land,
                                       seed=P * 7 + order)
                        key = (P, nthreads, order, phase, land)
                        assert got[2] == ref[2], key
                        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), key
                        assert abs(got[3] - ref[3]) <= 1e-12 * max(1.0, ref[3]), key


def test_wave_schedule_stops_like_the_sequential_loop():
    ix, iy, rho, u, v, _ = system(48, 40, seed=3)
    for tol in (1e-1, 1e-2, 1e-3):
        ref = run_seq(ix, iy, rho, u, v, 7.0, tol, 150)
        got = run_wave(ix, iy, rho, u, v, 7.0, tol, 150, 2, 16, 2, 1, 2)
        assert 1 < ref[2] < 150
        assert got[2] == ref[2]
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])


def test_unsupported_sizes_are_refused():
    ix, iy, rho, u, v, _ = system(5, 5, seed=1)
    assert run_wave(ix[:2], iy[:2], rho[:2], u[:2], v[:2], 7.0, 0.0, 1, 1, 4, 0, 0, 0)[2] == -1


@pytest.mark.parametrize("nx,ny", [(37, 29), (64, 48)])
def test_fp32_sequential_sweep_tracks_the_fp64_oracle(oracle_f64, nx, ny):
    """The kernel's arithmetic (fp32, system formed from I2wx, I2wy, rho_c on the fly) against the
    oracle's (fp64, stored Au, Av, Du, Dv, D): same sweep, differences at fp32 rounding level."""
    ix, iy, rho, u, v, x = system(nx, ny, seed=11)
    sysm = oracle_f64.hs_system(x["I1"], x["I2w"], x["I2wx"], x["I2wy"], x["u"], x["v"], 7.0)
    ou, ov, on, oerr = oracle_f64.hs_sor(*sysm, x["u"], x["v"], 7.0, tol=1e-3, maxiter=150)
    su, sv, sn, serr = run_seq(ix, iy, rho, u, v, 7.0, 1e-3, 150)
    assert sn == on
    assert np.abs(su - ou).max() < 2e-4 and np.abs(sv - ov).max() < 2e-4
    assert abs(serr - oerr) < 1e-5


# ---- the pipelined schedule (hs_sor_pipe.h) ------------------------------------------------------

@pytest.mark.parametrize("nx,ny", SHAPES + [(70, 23)])
def test_pipelined_schedule_equals_sequential_loop_fixed_count(nx, ny):
    ix, iy, rho, u, v, _ = system(nx, ny, seed=nx * 100 + ny)
    for maxiter in (1, 2, 7):
        ref = run_seq(ix, iy, rho, u, v, 7.0, 0.0, maxiter)
        for P in (0, 1, 3):
            for nthreads in sorted({1, 3, ny, ny + 5}):
                for order in range(3):
                    for phase in range(3):
                        for land in range(3):
                            got = run_pipe_wave(ix, iy, rho, u, v, 7.0, 0.0, maxiter, 4, P, nthreads, order, phase,
                                                land, seed=P * 7 + order)
                            key = (maxiter, P, nthreads, order, phase, land)
                            assert got[2] == ref[2] == maxiter, key
                            assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), key
                            assert abs(got[3] - ref[3]) <= 1e-12 * max(1.0, ref[3]), key
                            assert got[4] == 0, key


@pytest.mark.parametrize("nx,ny", [(12, 9), (37, 29), (64, 48), (23, 70), (70, 23)])
@pytest.mark.parametrize("K", [1, 4, 8])
def test_pipelined_schedule_stops_exactly_where_the_sequential_loop_stops(nx, ny, K):
    ix, iy, rho, u, v, _ = system(nx, ny, seed=7 * nx + ny)
    seen = set()
    for tol in (3e-1, 1e-1, 3e-2, 1e-2, 3e-3, 1e-3):
        ref = run_seq(ix, iy, rho, u, v, 7.0, tol, 150)
        if ref[2] in seen:
            continue
        seen.add(ref[2])
        for order, phase, land in ((0, 0, 0), (1, 2, 1), (2, 1, 2), (2, 0, 2)):
            got = run_pipe_wave(ix, iy, rho, u, v, 7.0, tol, 150, K, 1 + order, max(1, ny // 2), order, phase, land)
            assert got[2] == ref[2], (tol, order, got[2], ref[2])
            assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), (tol, order)
            assert abs(got[3] - ref[3]) <= 1e-12 * max(1.0, ref[3])
    assert len(seen) >= 3


# ---- two columns per thread-step (hs_sor_pairs.h; not yet a kernel) ------------------------------

@pytest.mark.parametrize("nx,ny", SHAPES + [(70, 23), (6, 5), (7, 6)])
def test_pairs_schedule_equals_sequential_loop_fixed_count(nx, ny):
    ix, iy, rho, u, v, _ = system(nx, ny, seed=nx * 100 + ny)
    for maxiter in (1, 2, 7):
        ref = run_seq(ix, iy, rho, u, v, 7.0, 0.0, maxiter)
        for P in (0, 1, 3):
            for nthreads in sorted({1, 3, ny, ny + 5}):
                for order in range(3):
                    for phase in range(3):
                        for land in range(3):
                            got = run_pipe_wave(ix, iy, rho, u, v, 7.0, 0.0, maxiter, 4, P, nthreads, order, phase,
                                                land, seed=P * 7 + order, pairs=True)
                            key = (maxiter, P, nthreads, order, phase, land)
                            assert got[2] == ref[2] == maxiter, key
                            assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), key
                            assert abs(got[3] - ref[3]) <= 1e-12 * max(1.0, ref[3]), key
                            assert got[4] == 0, key


@pytest.mark.parametrize("nx,ny", [(12, 9), (37, 29), (64, 48), (23, 70), (70, 23)])
@pytest.mark.parametrize("K", [1, 4, 8])
def test_pairs_schedule_stops_exactly_where_the_sequential_loop_stops(nx, ny, K):
    ix, iy, rho, u, v, _ = system(nx, ny, seed=7 * nx + ny)
    seen = set()
    for tol in (3e-1, 1e-1, 3e-2, 1e-2, 3e-3, 1e-3):
        ref = run_seq(ix, iy, rho, u, v, 7.0, tol, 150)
        if ref[2] in seen:
            continue
        seen.add(ref[2])
        for order, phase, land in ((0, 0, 0), (1, 2, 1), (2, 1, 2), (2, 0, 2)):
            got = run_pipe_wave(ix, iy, rho, u, v, 7.0, tol, 150, K, 1 + order, max(1, ny // 2), order, phase, land,
                                pairs=True)
            assert got[2] == ref[2], (tol, order, got[2], ref[2])
            assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), (tol, order)
            assert abs(got[3] - ref[3]) <= 1e-12 * max(1.0, ref[3])
    assert len(seen) >= 3
