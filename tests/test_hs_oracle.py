"""Pins the Horn-Schunck part of the CPU oracle (oracle/tvl1_oracle.c, section (d)) to the reference's
src/horn_schunck_pyramidal.cpp:

1. against tests/golden/hs_reference_vectors.npz -- outputs of the unmodified reference objects run
   with ONE thread (tests/golden/make_golden.py), bit-for-bit, double and float builds;
2. against oracle/_ref directly when that library is present.

One thread, because the reference's SOR sweep updates u and v in place inside an OpenMP parallel-for
(src/horn_schunck_pyramidal.cpp:148-158): with several threads its own result depends on scheduling
(test_reference_sweep_is_only_defined_for_one_thread shows it).
"""
import os

import numpy as np
import pytest

import _cases
from oracle.loader import CpuTvl1, available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hs_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "hs_reference_vectors.npz"))


@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("name", sorted(_cases.HS_CASES))
def test_hs_solver_matches_golden(hs_golden, oracle_f64, oracle_f32, tag, name):
    cpu = oracle_f64 if tag == "f64" else oracle_f32
    u, v, iters, errs = _cases.run_hs_case(cpu, _cases.HS_CASES[name])
    pre = "%s/%s/" % (tag, name)
    assert np.array_equal(iters, hs_golden[pre + "iters"])
    assert np.array_equal(u, hs_golden[pre + "u"])
    assert np.array_equal(v, hs_golden[pre + "v"])
    # the reference prints the error with %g (6 significant digits): :233-235
    assert np.allclose(errs, hs_golden[pre + "errs"], rtol=1e-5)


def test_hs_golden_covers_both_stopping_rules(hs_golden):
    # :143  while (error > TOL && niter < maxiter)
    assert hs_golden["f64/hs_64x48/iters"].max() == 60            # capped by maxiter
    assert hs_golden["f64/hs_96x64_tol/iters"].max() < 150        # stopped by TOL
    assert hs_golden["f64/hs_96x64_tol/errs"].max() <= 2e-2


def _sequential_sweep(Au, Av, Du, Dv, D, u, v, al):
    """Pure-Python statement of ONE sweep (src/horn_schunck_pyramidal.cpp:144-230) written from the
    reference's explicit neighbour lists, independent of the clamped-index form the C oracle uses."""
    ny, nx = u.shape
    u, v = u.ravel().copy(), v.ravel().copy()
    Au, Av, Du, Dv, D = (a.ravel() for a in (Au, Av, Du, Dv, D))
    w = 1.9

    def sor(p, p1, p2, p3, p4, p5, p6, p7, p8):
        ula = 1. / 12. * (u[p1] + u[p2] + u[p3] + u[p4]) + 1. / 6. * (u[p5] + u[p6] + u[p7] + u[p8])
        vla = 1. / 12. * (v[p1] + v[p2] + v[p3] + v[p4]) + 1. / 6. * (v[p5] + v[p6] + v[p7] + v[p8])
        uk, vk = u[p], v[p]
        u[p] = (1.0 - w) * uk + w * (Au[p] - D[p] * v[p] + al * ula) / Du[p]
        v[p] = (1.0 - w) * vk + w * (Av[p] - D[p] * u[p] + al * vla) / Dv[p]
        return (u[p] - uk) * (u[p] - uk) + (v[p] - vk) * (v[p] - vk)

    e = 0.0
    for i in range(1, ny - 1):
        for j in range(1, nx - 1):
            k = i * nx + j
            e += sor(k, k - nx - 1, k - nx + 1, k + nx - 1, k + nx + 1, k - nx, k - 1, k + nx, k + 1)
    for j in range(1, nx - 1):
        k = j
        e += sor(k, k - 1, k + 1, k + nx - 1, k + nx + 1, k, k - 1, k + nx, k + 1)
        k = (ny - 1) * nx + j
        e += sor(k, k - nx - 1, k - nx + 1, k - 1, k + 1, k - nx, k - 1, k, k + 1)
    for i in range(1, ny - 1):
        k = i * nx
        e += sor(k, k - nx, k - nx + 1, k + nx, k + nx + 1, k - nx, k, k + nx, k + 1)
        k = (i + 1) * nx - 1
        e += sor(k, k - nx - 1, k - nx, k + nx - 1, k + nx, k - nx, k - 1, k + nx, k)
    e += sor(0, 0, 1, nx, nx + 1, 0, 0, nx, 1)
    k = nx - 1
    e += sor(k, k - 1, k, k + nx - 1, k + nx, k, k - 1, k + nx, k)
    k = (ny - 1) * nx
    e += sor(k, k - nx, k - nx + 1, k, k + 1, k - nx, k, k, k + 1)
    k = ny * nx - 1
    e += sor(k, k - 1, k, k - nx - 1, k - nx, k - nx, k - 1, k, k)
    return u.reshape(ny, nx), v.reshape(ny, nx), np.sqrt(e / (nx * ny))


@pytest.mark.parametrize("shape", [(29, 37), (3, 3), (3, 9), (8, 3), (4, 5)])
def test_hs_sor_sweep_order(oracle_f64, shape):
    """The C oracle's clamped-neighbour sweep is the reference's explicit one, corners included."""
    ny, nx = shape
    x = _cases.hs_sor_inputs(nx, ny)
    sysm = oracle_f64.hs_system(x["I1"], x["I2w"], x["I2wx"], x["I2wy"], x["u"], x["v"], 7.0)
    u, v, n, err = oracle_f64.hs_sor(*sysm, x["u"], x["v"], 7.0, tol=0.0, maxiter=1)
    ru, rv, rerr = _sequential_sweep(*sysm, x["u"], x["v"], 49.0)
    assert n == 1
    assert np.array_equal(u, ru) and np.array_equal(v, rv)
    assert err == rerr


@pytest.mark.skipif(not (available("reference", np.float64) and available("reference", np.float32)),
                    reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_hs_port_equals_compiled_reference(dt):
    P, R = CpuTvl1("port", dt), CpuTvl1("reference", dt)
    I1, I2 = _cases.synth.make_pair(80, 56, seed=11, scale=0.5)
    a = P.hs_multiscale(I1, I2, nscales=3, warps=3, maxiter=50)
    b = R.hs_multiscale(I1, I2, nscales=3, warps=3, maxiter=50)
    assert np.array_equal(a[2], b[2])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    z = np.zeros_like(I1)
    a = P.hs_single_scale(I1, I2, z, z, warps=2, maxiter=25)
    b = R.hs_single_scale(I1, I2, z, z, warps=2, maxiter=25)
    assert np.array_equal(a[2], b[2])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.skipif(not available("reference", np.float64), reason="oracle/_ref not built")
def test_reference_sweep_is_only_defined_for_one_thread():
    """Documents why parity is pinned at one thread: with several OpenMP threads the reference's
    in-place sweep reads neighbours other threads are updating, and its result moves away from the
    one-thread result (by far more than rounding)."""
    R = CpuTvl1("reference", np.float64)
    ncpu = os.cpu_count() or 1          # (omp_get_max_threads is process-global: other tests set it to 1)
    if ncpu < 2:
        pytest.skip("one hardware thread")
    I1, I2 = _cases.synth.make_pair(160, 120, seed=3, scale=0.5)
    kw = dict(nscales=2, warps=2, maxiter=30)
    one = R.hs_multiscale(I1, I2, threads=1, **kw)
    many = R.hs_multiscale(I1, I2, threads=min(8, ncpu), **kw)
    again = R.hs_multiscale(I1, I2, threads=1, **kw)
    assert np.array_equal(one[0], again[0])
    assert np.abs(one[0] - many[0]).max() > 1e-9
