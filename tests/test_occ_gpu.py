"""GPU parity tests of the TV-L1 + occlusions path (include/occ_b200.h) through the C ABI.

The solver computes in IEEE fp64 with the reference's association order, so the bar is the integer bar:
flow, occlusion map, dual variables and iteration counts BIT-IDENTICAL to the reference --

* the box relaxation (wavefront schedule of the lexicographic Gauss-Seidel pass) and the median filter
  against the C restatement (and, where oracle/_ref travelled, the compiled reference objects);
* the solver against the committed golden vectors of the unmodified reference
  (tests/golden/occ_reference_vectors.npz) and against the oracle on fresh cases;
* batch = individual solves; single-level entry point; the reference's mangled C++ symbols.
"""
import ctypes as C
import os

import numpy as np
import pytest

import _cases
import optical_flow_1_b200 as pkg
from oracle.loader import CpuOcc, occ_available

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gpu():
    g = pkg.TVL1Occ(device=0)
    yield g
    g.close()


@pytest.fixture(scope="module")
def port():
    p = CpuOcc("port", np.float64)
    p.set_threads(1)
    return p


@pytest.fixture(scope="module")
def occ_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "occ_reference_vectors.npz"))


def box_system(nx, ny, seed):
    rs = np.random.RandomState(seed)
    return dict(f=rs.uniform(-5, 5, (ny, nx)), u=rs.uniform(-3, 3, (ny, nx)), g=rs.uniform(0.2, 1, (ny, nx)),
                p1=rs.uniform(-1, 1, (ny, nx)), p2=rs.uniform(-1, 1, (ny, nx)))


@pytest.mark.parametrize("nx,ny", [(2, 2), (3, 3), (9, 7), (7, 9), (33, 20), (64, 5), (5, 64), (131, 70), (96, 80)])
def test_box_relaxation_is_the_sequential_sweep_bitwise(gpu, port, nx, ny):
    """Every corner / side / interior system, warm-started duals, 10 sweeps then 3 more on the result."""
    s = box_system(nx, ny, nx * 31 + ny)
    a = port.rof_box(s["u"], s["f"], s["p1"], s["p2"], s["g"], 0.3, 1.25, 10)
    b = gpu.rof_box(s["u"], s["f"], s["p1"], s["p2"], s["g"], 0.3, 1.25, 10)
    for x, y in zip(a, b):
        assert np.array_equal(x, y), np.abs(x - y).max()
    a = port.rof_box(a[0], s["f"], a[1], a[2], s["g"], 0.3, 1.25, 3)
    b = gpu.rof_box(b[0], s["f"], b[1], b[2], s["g"], 0.3, 1.25, 3)
    for x, y in zip(a, b):
        assert np.array_equal(x, y), np.abs(x - y).max()


def test_box_relaxation_rows_beyond_one_thread_each(gpu, port):
    """More rows than threads of a CTA (1024): a thread owns rows tid, tid + blockDim, ..."""
    s = box_system(24, 2100, 5)
    a = port.rof_box(s["u"], s["f"], s["p1"], s["p2"], s["g"], 0.3, 1.25, 2)
    b = gpu.rof_box(s["u"], s["f"], s["p1"], s["p2"], s["g"], 0.3, 1.25, 2)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.mark.skipif(not occ_available("reference", np.float64), reason="oracle/_ref not built")
def test_box_relaxation_equals_the_compiled_reference(gpu):
    s = box_system(160, 120, 77)
    R = CpuOcc("reference")
    a = R.rof_box(s["u"], s["f"], s["p1"], s["p2"], s["g"], 0.3, 1.25, 10)
    b = gpu.rof_box(s["u"], s["f"], s["p1"], s["p2"], s["g"], 0.3, 1.25, 10)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("shape", [(11, 13), (1, 9), (9, 1), (2, 2), (120, 67)])
def test_median_filter(gpu, port, shape):
    rs = np.random.RandomState(3)
    a = rs.uniform(-1, 1, shape)
    a[rs.uniform(0, 1, shape) < 0.2] = 0.0            # ties, and zeros of both signs
    a[rs.uniform(0, 1, shape) < 0.1] = -0.0
    m, r = gpu.median3(a), port.median3(a)
    assert np.array_equal(m, r) and np.array_equal(np.signbit(m), np.signbit(r))


@pytest.mark.parametrize("name", sorted(_cases.OCC_CASES))
def test_solver_equals_golden_bit_for_bit(gpu, occ_golden, name):
    case = _cases.OCC_CASES[name]
    I_1, I0, I1 = _cases.occ_inputs(case)
    u1, u2, chi, iters, _ = gpu.Dual_TVL1_optic_flow_multiscale(I_1, I0, I1, None, **case["kw"])
    assert np.array_equal(iters, occ_golden["f64/%s/iters" % name]), (iters, occ_golden["f64/%s/iters" % name])
    assert np.array_equal(u1, occ_golden["f64/%s/u1" % name]), np.abs(u1 - occ_golden["f64/%s/u1" % name]).max()
    assert np.array_equal(u2, occ_golden["f64/%s/u2" % name])
    assert np.array_equal(chi.astype(np.uint8), occ_golden["f64/%s/chi" % name])
    assert set(np.unique(chi)) <= {0.0, 1.0}
    st = gpu.stats()
    assert st["kernel_launches"] > 0 and st["box_sweeps"] > 0


@pytest.mark.parametrize("env", [{"OCC_GS_WAVE": "0"}, {"OCC_GS_COEF": "1"}, {"OCC_CHI_TB": "1"}, {"OCC_CHI_MARCH": "1"},
                                 {"OCC_CHI_FUSED": "0"}],
                         ids=["gs_row_major", "gs_coefficient_kernel", "chi_temporal_blocking", "chi_row_march", "chi_two_kernels"])
def test_kernel_variants_give_the_same_bits(occ_golden, env, monkeypatch):
    """The A/B variants kept in the library (row-major Gauss-Seidel pass, coefficient kernel, temporally blocked
    row-marching and two-kernel occlusion-map iteration) against the golden vectors: every variant is the same arithmetic."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    g = pkg.TVL1Occ(device=0)
    for name in ("occ_96x80_tight", "occ_61x47_z07"):
        case = _cases.OCC_CASES[name]
        I_1, I0, I1 = _cases.occ_inputs(case)
        u1, u2, chi, iters, _ = g.Dual_TVL1_optic_flow_multiscale(I_1, I0, I1, None, **case["kw"])
        assert np.array_equal(iters, occ_golden["f64/%s/iters" % name])
        assert np.array_equal(u1, occ_golden["f64/%s/u1" % name]) and np.array_equal(u2, occ_golden["f64/%s/u2" % name])
        assert np.array_equal(chi.astype(np.uint8), occ_golden["f64/%s/chi" % name])
    g.close()


def fresh_case(nx, ny, seed, **kw):
    p = dict(lam=0.15, alpha=0.01, beta=0.15, theta=0.3, nscales=3, zfactor=0.5, warps=2, eps=0.01)
    p.update(kw)
    return dict(nx=nx, ny=ny, seed=seed, scale=0.6, kw=p)


@pytest.mark.parametrize("case", [fresh_case(120, 90, 11), fresh_case(160, 120, 5, nscales=4, warps=3, eps=0.003),
                                  fresh_case(97, 61, 21, zfactor=0.6, nscales=3),
                                  fresh_case(320, 240, 3, nscales=4, warps=2)],
                         ids=["120x90", "160x120_tight", "97x61_z06", "320x240"])
def test_solver_equals_oracle_on_fresh_cases(gpu, port, case):
    I_1, I0, I1 = _cases.occ_inputs(case)
    r = port.multiscale(I_1, I0, I1, None, **case["kw"])
    g = gpu.Dual_TVL1_optic_flow_multiscale(I_1, I0, I1, None, **case["kw"])
    assert np.array_equal(g[3], r[3]), (g[3], r[3])
    for k in range(3):
        assert np.array_equal(g[k], r[k]), (k, np.abs(g[k] - r[k]).max())
    assert np.allclose(g[4], r[4], rtol=1e-9, atol=1e-300)     # errors: same sum, different (fixed) order


def test_separate_weight_image(gpu, port):
    """filtI0 given explicitly (the CLI's optional fourth image): its own pyramid, its own g."""
    case = fresh_case(112, 84, 8)
    I_1, I0, I1 = _cases.occ_inputs(case)
    rs = np.random.RandomState(1)
    filt = I0.astype(np.float64) + rs.uniform(-4, 4, I0.shape)
    r = port.multiscale(I_1, I0, I1, filt, **case["kw"])
    g = gpu.Dual_TVL1_optic_flow_multiscale(I_1, I0, I1, filt, **case["kw"])
    assert np.array_equal(g[3], r[3])
    for k in range(3):
        assert np.array_equal(g[k], r[k])


def test_batch_equals_individual_solves(gpu):
    """Triples advance in lock-step and stop on their own criteria: same bits as one by one."""
    cases = [fresh_case(96, 72, s, eps=e) for s, e in ((1, 0.01), (2, 0.002), (3, 0.05))]
    trip = [_cases.occ_inputs(c) for c in cases]
    kw = dict(cases[0]["kw"])
    kw["eps"] = 0.004
    one = [gpu.Dual_TVL1_optic_flow_multiscale(*t, None, **kw) for t in trip]
    I_1, I0, I1 = (np.stack([t[k] for t in trip]) for k in range(3))
    u1, u2, chi, iters, _ = gpu.Dual_TVL1_optic_flow_multiscale(I_1, I0, I1, None, **kw)
    for b, o in enumerate(one):
        assert np.array_equal(iters[b], o[3])
        assert np.array_equal(u1[b], o[0]) and np.array_equal(u2[b], o[1]) and np.array_equal(chi[b], o[2])


def test_single_level_entry(gpu, port):
    """Dual_TVL1_optic_flow (7 planes): one level on the images as given, from the given flow and map,
    against the oracle's single-level function."""
    case = fresh_case(80, 64, 4, nscales=1)
    I_1, I0, I1 = (x.astype(np.float64) for x in _cases.occ_inputs(case))
    rs = np.random.RandomState(2)
    u0, v0 = rs.uniform(-1, 1, I0.shape), rs.uniform(-1, 1, I0.shape)
    c0 = rs.uniform(0, 1, I0.shape)
    g = gpu.Dual_TVL1_optic_flow(I_1, I0, I1, None, u0, v0, c0, warps=2, eps=0.01)
    r = port.single_scale(I_1, I0, I1, None, u0, v0, c0, warps=2, eps=0.01)
    assert np.array_equal(g[3], r[3])
    for k in range(3):
        assert np.array_equal(g[k], r[k])
    assert 0 < g[2].min() or g[2].max() <= 1          # chi is not thresholded here


def test_mangled_reference_symbols(gpu, port):
    """The reference's C++ entry point (src/tvl1occflow.h:111-129, ofpix_t = double) exported by the library."""
    lib = C.CDLL(pkg.library_path())
    fn = getattr(lib, "_Z31Dual_TVL1_optic_flow_multiscalePdS_S_S_S_S_S_iiddddididb")
    fn.restype = None
    case = _cases.OCC_CASES["occ_64x48"]
    I_1, I0, I1 = (np.ascontiguousarray(x, np.float64) for x in _cases.occ_inputs(case))
    u1, u2, chi = (np.full(I0.shape, 7.0) for _ in range(3))       # contents on entry are ignored
    kw = case["kw"]
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    fn(p(I_1), p(I0), p(I1), p(I0), p(u1), p(u2), p(chi), C.c_int(case["nx"]), C.c_int(case["ny"]),
       C.c_double(kw["lam"]), C.c_double(kw["alpha"]), C.c_double(kw["beta"]), C.c_double(kw["theta"]),
       C.c_int(kw["nscales"]), C.c_double(kw["zfactor"]), C.c_int(kw["warps"]), C.c_double(kw["eps"]), C.c_bool(False))
    r = _cases.run_occ_case(port, case)
    assert np.array_equal(u1, r[0]) and np.array_equal(u2, r[1]) and np.array_equal(chi, r[2])


def test_sigma_too_large_is_reported(gpu):
    I = np.zeros((8, 4))
    with pytest.raises(pkg.OccError) as e:
        gpu.Dual_TVL1_optic_flow_multiscale(I, I, I, None, nscales=1, warps=1)
    assert e.value.code == 2 and "sigma too large" in str(e.value)
