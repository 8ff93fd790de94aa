"""CPU tests of the TV-L1 + occlusions oracle (oracle/tvl1_oracle.c section (e)): the C
restatement against the committed golden vectors of the unmodified reference (built with the
zero-filling new[] of oracle/occ_ref_shim.cpp -- the one defined reading of the reference's
uninitialised eta1/eta2), and, when oracle/_ref travelled, against the compiled reference itself."""
import os

import numpy as np
import pytest

import _cases
from oracle.loader import CpuOcc, occ_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def occ_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "occ_reference_vectors.npz"))


@pytest.mark.parametrize("name", sorted(_cases.OCC_CASES))
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_port_equals_golden_bit_for_bit(occ_golden, name, dt):
    tag = "f64" if dt == np.float64 else "f32"
    P = CpuOcc("port", dt)
    P.set_threads(1)
    u1, u2, chi, iters, _ = _cases.run_occ_case(P, _cases.OCC_CASES[name])
    assert np.array_equal(iters, occ_golden["%s/%s/iters" % (tag, name)])
    assert np.array_equal(u1, occ_golden["%s/%s/u1" % (tag, name)])
    assert np.array_equal(u2, occ_golden["%s/%s/u2" % (tag, name)])
    assert np.array_equal(chi.astype(np.uint8), occ_golden["%s/%s/chi" % (tag, name)])
    assert set(np.unique(chi)) <= {0.0, 1.0}            # thresholded at THR_CHI (src/tvl1occflow.cpp:459)


def test_golden_cases_exercise_the_loops(occ_golden):
    """The fixtures are not degenerate: some warp steps run several outer iterations, the occlusion map is
    neither empty nor full, the flow has the synthetic motion's sign."""
    it = occ_golden["f64/occ_96x80_tight/iters"]
    assert it.max() > 1 and it.min() >= 1 and it.max() <= 20
    chi = occ_golden["f64/occ_96x80_tight/chi"]
    assert 0 < chi.sum() < chi.size
    assert np.median(occ_golden["f64/occ_96x80_tight/u1"]) > 0.3


@pytest.mark.skipif(not occ_available("reference", np.float64), reason="oracle/_ref not built")
@pytest.mark.parametrize("shape", [(9, 7), (33, 20), (3, 3), (2, 2), (64, 5)])
def test_box_relaxation_and_median_equal_the_compiled_reference(shape):
    """Scalar_ROF_BoxCellCentered (every corner / side / interior system, warm-started duals) and
    me_median_filtering, restated on compact arrays, against the reference objects: identical bits."""
    nx, ny = shape
    rs = np.random.RandomState(nx * 31 + ny)
    f, u0 = rs.uniform(-5, 5, (ny, nx)), rs.uniform(-3, 3, (ny, nx))
    g = rs.uniform(0.2, 1, (ny, nx))
    p1, p2 = rs.uniform(-1, 1, (ny, nx)), rs.uniform(-1, 1, (ny, nx))
    R, P = CpuOcc("reference"), CpuOcc("port")
    a = R.rof_box(u0, f, p1, p2, g, 0.3, 1.25, 10)
    b = P.rof_box(u0, f, p1, p2, g, 0.3, 1.25, 10)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    a = R.rof_box(a[0], f, a[1], a[2], g, 0.3, 1.25, 3)      # second call: duals carried over
    b = P.rof_box(b[0], f, b[1], b[2], g, 0.3, 1.25, 3)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    m = rs.uniform(-3, 3, (ny, nx))
    assert np.array_equal(R.median3(m), P.median3(m))


@pytest.mark.skipif(not occ_available("reference", np.float64), reason="oracle/_ref not built")
def test_port_equals_the_compiled_reference_on_a_fresh_case():
    """A case that is not among the fixtures, fp64: flows, occlusion map and iteration counts identical."""
    case = dict(nx=120, ny=90, seed=11, scale=0.6,
                kw=dict(lam=0.15, alpha=0.01, beta=0.15, theta=0.3, nscales=3, zfactor=0.5, warps=2, eps=0.01))
    R, P = CpuOcc("reference"), CpuOcc("port")
    a = _cases.run_occ_case(R, case)
    b = _cases.run_occ_case(P, case)
    assert np.array_equal(a[3], b[3])
    for k in range(3):
        assert np.array_equal(a[k], b[k])


def test_median_is_the_middle_of_nine():
    P = CpuOcc("port")
    rs = np.random.RandomState(3)
    a = rs.uniform(-1, 1, (11, 13))
    m = P.median3(a)
    pad = np.pad(a, 1, mode="symmetric")            # x<0 -> -x-1, x>=n -> 2n-x-1
    ref = np.median(np.stack([pad[i:i + 11, j:j + 13] for i in range(3) for j in range(3)]), axis=0)
    assert np.array_equal(m, ref)
