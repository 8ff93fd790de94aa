"""Pins the CPU oracle (oracle/tvl1_oracle.c) to the reference.

1. against tests/golden/reference_vectors.npz -- outputs of the unmodified reference objects
   (tests/golden/make_golden.py), bit-for-bit, double and float builds;
2. against oracle/_ref directly when that library is present (build container and GPU box).
"""
import numpy as np
import pytest

import _cases
from oracle import loader
from oracle.loader import CpuTvl1, available


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_functions_match_golden(golden, oracle_f64, oracle_f32, tag):
    cpu = oracle_f64 if tag == "f64" else oracle_f32
    got = _cases.run_function_cases(cpu)
    for k, v in got.items():
        ref = golden["%s/%s" % (tag, k)]
        assert v.shape == ref.shape, k
        assert np.array_equal(v, ref), "%s/%s differs: max |d| = %g" % (tag, k, np.abs(v - ref).max())


@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("name", sorted(_cases.SOLVER_CASES))
def test_solver_matches_golden(golden, oracle_f64, oracle_f32, tag, name):
    cpu = oracle_f64 if tag == "f64" else oracle_f32
    u1, u2, iters, errs = _cases.run_solver_case(cpu, _cases.SOLVER_CASES[name])
    pre = "%s/%s/" % (tag, name)
    assert np.array_equal(iters, golden[pre + "iters"])
    assert np.array_equal(u1, golden[pre + "u1"])
    assert np.array_equal(u2, golden[pre + "u2"])
    # the reference prints errors with %f (6 decimals): src/tvl1flow.cpp:185-187
    assert np.allclose(errs, golden[pre + "errs"], atol=5.1e-7)


def test_golden_has_iteration_cap_case(golden):
    # eps = 0.0003 on the 96x64 case drives the coarse warps into MAX_ITERATIONS (tvl1flow.cpp:22,113)
    assert golden["f64/ms_96x64_cap/iters"].max() == 300
    assert golden["f64/ms_64x48/iters"].min() >= 1


def test_warp_validity_box(oracle_f64):
    """border_out = true returns exactly 0 unless 1 <= uu < nx-2 and 1 <= vv < ny-2
    (src/bicubic_interpolation.cpp:165-215)."""
    x = _cases.func_inputs()
    I, u, v = x["I"], x["u"], x["v"]
    ny, nx = I.shape
    w = oracle_f64.warp(I, u, v, True)
    jj, ii = np.meshgrid(np.arange(nx), np.arange(ny))
    uu, vv = jj + u, ii + v
    inside = (uu >= 1) & (uu < nx - 2) & (vv >= 1) & (vv < ny - 2)
    assert np.all(w[~inside] == 0.0)
    assert np.all(w[inside] != 0.0)


def test_zoom_out_half_is_blur_then_decimate(oracle_f64):
    """factor 0.5 samples land on integers, where the cubic returns v[1] exactly
    (src/zoom.cpp:67-75, src/bicubic_interpolation.cpp:114-122)."""
    I = _cases.func_inputs()["I"]
    blurred = oracle_f64.gaussian(I, _cases.SIGMA_ZOOM_HALF)
    z = oracle_f64.zoom_out(I, 0.5)
    assert np.array_equal(z, blurred[0:2 * z.shape[0]:2, 0:2 * z.shape[1]:2])


def test_gaussian_sigma_too_large(oracle_f64):
    # src/operators.cpp:520-522: window (int)(5*sigma)+1 wider than the image -> exception
    with pytest.raises(RuntimeError):
        oracle_f64.gaussian(np.ones((8, 4)), 0.8)


def test_eps_zero_runs_to_cap(oracle_f64):
    I0, I1 = _cases.synth.make_pair(32, 24, seed=3, scale=0.3)
    z = np.zeros_like(I0)
    _, _, iters, _ = oracle_f64.single_scale(I0, I1, z, z, warps=2, eps=0.0)
    assert iters.tolist() == [300, 300]


@pytest.mark.skipif(not (available("reference", np.float64) and available("reference", np.float32)),
                    reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_port_equals_compiled_reference(dt):
    P, R = CpuTvl1("port", dt), CpuTvl1("reference", dt)
    P.set_threads(1)
    R.set_threads(1)
    a, b = _cases.run_function_cases(P), _cases.run_function_cases(R)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    I0, I1 = _cases.synth.make_pair(80, 56, seed=11, scale=0.5)
    ra = P.multiscale(I0, I1, nscales=3, warps=3)
    rb = R.multiscale(I0, I1, nscales=3, warps=3)
    assert np.array_equal(ra[2], rb[2])
    assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])


@pytest.mark.skipif(not (loader.upstream_c99_available() and available("reference", np.float32)),
                    reason="oracle/_ref not built (needs /root/reference at build time)")
def test_upstream_c99_library_is_the_same_algorithm():
    """The IPOL C99 library the reference derives from (3rdparty/tvl1flow_3/tvl1flow_lib.c, float, C
    linkage) against the reference's C++ sources built with ofpix_t = float: same pyramid, same
    stopping behaviour.  This is what lets one CUDA path serve both sets of entry points."""
    import ctypes as C
    up = loader.c99_signature(C.CDLL(loader.UPSTREAM_C99))
    R = CpuTvl1("reference", np.float32)
    for (nx, ny, ns, seed) in [(160, 120, 3, 1234), (211, 173, 4, 7)]:
        I0, I1 = _cases.synth.make_pair(nx, ny, seed=seed)
        u1, u2 = loader.c99_multiscale(up, I0, I1, nscales=ns)
        r1, r2, _, _ = R.multiscale(I0.astype(np.float32), I1.astype(np.float32), nscales=ns)
        d = np.abs(np.stack([u1 - r1, u2 - r2]))
        assert d.mean() < 1e-4 and d.max() < 1e-2, (d.mean(), d.max())
