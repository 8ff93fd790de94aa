"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs, against the committed golden vectors of the unmodified reference, and -- at
BASELINE.json's full sizes -- through size-independent properties.

Tolerances (BASELINE.json north_star; the CUDA path computes in fp32, the reference in fp64):
  flow: mean |d| <= 1e-3 px and max |d| <= 1e-2 px, iteration counts identical per (level, warp).
Per-kernel hooks are compared with fp32-rounding-sized bounds stated at each test.
"""
import os

import numpy as np
import pytest

import _cases
import optical_flow_1_b200 as pkg
from oracle.loader import CpuTvl1, available

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MEAN_TOL = 1e-3
MAX_TOL = 1e-2


@pytest.fixture(scope="module")
def gpu():
    g = pkg.TVL1(device=0)
    yield g
    g.close()


def flow_diff(u1, u2, r1, r2):
    d = np.concatenate([np.abs(u1.astype(np.float64) - r1).ravel(),
                        np.abs(u2.astype(np.float64) - r2).ravel()])
    return d.mean(), d.max()


def assert_flow_close(u1, u2, r1, r2, what=""):
    mean, mx = flow_diff(u1, u2, r1, r2)
    assert mean <= MEAN_TOL and mx <= MAX_TOL, "%s mean|d|=%g max|d|=%g" % (what, mean, mx)


# ---- per-kernel hooks vs the oracle (fp64 port) -------------------------------------------------

def test_normalize(gpu, oracle_f64):
    x = _cases.func_inputs()
    a0, a1 = gpu.image_normalization_2(x["I"], x["J"])
    r0, r1 = oracle_f64.normalize(x["I"].astype(np.float32), x["J"].astype(np.float32))
    # values in [0,255]: a few fp32 ulps (255 * 6e-8 ~ 1.5e-5)
    assert np.abs(a0 - r0).max() < 1e-4 and np.abs(a1 - r1).max() < 1e-4
    # constant images are copied unchanged (src/utils.cpp:319-325)
    c = np.full((9, 12), 7.25, np.float32)
    b0, b1 = gpu.image_normalization_2(c, c)
    assert np.array_equal(b0, c) and np.array_equal(b1, c)


@pytest.mark.parametrize("sigma", [_cases.SIGMA_PRE, _cases.SIGMA_ZOOM_HALF, 1.7])
@pytest.mark.parametrize("shape", [_cases.FUNC_SHAPE, (70, 131), (9, 7)])
def test_gaussian(gpu, oracle_f64, sigma, shape):
    I = np.random.RandomState(3).uniform(0, 255, shape).astype(np.float32)
    if (int)(5 * sigma) + 1 > shape[1]:
        with pytest.raises(pkg.TVL1Error) as e:
            gpu.gaussian(I, sigma)
        assert e.value.code == 2     # the reference throws here (src/operators.cpp:520-522)
        with pytest.raises(RuntimeError):
            oracle_f64.gaussian(I, sigma)
        return
    if (int)(5 * sigma) + 1 > shape[0]:
        pytest.skip("window taller than the image: undefined behaviour in the reference")
    got = gpu.gaussian(I, sigma)
    ref = oracle_f64.gaussian(I, sigma)
    assert np.abs(got - ref).max() < 2e-4     # ~20 fp32 roundings of values <= 255


@pytest.mark.parametrize("factor", [0.5, 0.7, 0.3])
@pytest.mark.parametrize("shape", [_cases.FUNC_SHAPE, (109, 218)])
def test_zoom_out(gpu, oracle_f64, factor, shape):
    I = np.random.RandomState(4).uniform(0, 255, shape).astype(np.float32)
    got = gpu.zoom_out(I, factor)
    ref = oracle_f64.zoom_out(I, factor)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 5e-4


@pytest.mark.parametrize("shape", [(1080, 1920), (436, 1024), (55, 109), (47, 61), (70, 131), (33, 240), (130, 122)])
def test_gaussian_kernel_variants_agree_bitwise(shape, monkeypatch):
    """The marching blur with shuffled row inputs (k_gauss_shfl, the default) against the one that loads its
    own inputs (k_gauss_march): blur, decimating blur (zoom_out at 0.5) and the normalising blur inside a
    solve -- identical bits, at sizes with ragged last groups, odd widths and fewer columns than a warp."""
    rs = np.random.RandomState(shape[0] + shape[1])
    I = rs.uniform(0, 255, shape).astype(np.float32)
    J = rs.uniform(10, 200, shape).astype(np.float32)
    out = []
    for flag in ("1", "0"):
        monkeypatch.setenv("TVL1_GAUSS_SHFL", flag)
        g = pkg.TVL1(device=0)
        res = [g.gaussian(I, _cases.SIGMA_PRE), g.zoom_out(I, 0.5)]
        if shape[0] >= 47:
            u1, u2, it, _ = g.Dual_TVL1_optic_flow_multiscale(I, J, nscales=2, warps=1, eps=0.05)
            res += [u1, u2, it]
        out.append(res)
        g.close()
    for a, b in zip(*out):
        assert np.array_equal(a, b)


def test_zoom_in(gpu, oracle_f64):
    I = np.random.RandomState(5).uniform(-8, 8, (27, 35)).astype(np.float32)
    for (nxx, nyy) in [(70, 54), (71, 53), (69, 55), (50, 38)]:
        got = gpu.zoom_in(I, nxx, nyy, scale=2.0)
        ref = oracle_f64.zoom_in(I, nxx, nyy) * 2.0
        assert np.abs(got - ref).max() < 2e-5


@pytest.mark.parametrize("shape", [_cases.FUNC_SHAPE, (64, 128), (50, 77)])
def test_warp_precompute(gpu, oracle_f64, shape):
    rs = np.random.RandomState(6)
    ny, nx = shape
    yy, xx = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    I1 = (128 + 60 * np.sin(0.21 * xx + 0.13 * yy) + 40 * np.cos(0.17 * yy - 0.05 * xx)).astype(np.float32)
    I0 = (I1 + rs.uniform(-5, 5, shape)).astype(np.float32)
    # flows that cross the validity box on every side, fractional parts spread over [0,1)
    u1 = rs.uniform(-6, 6, shape).astype(np.float32)
    u2 = rs.uniform(-6, 6, shape).astype(np.float32)
    u1[0, :] = 0.0
    u2[:, 0] = 0.0
    got = gpu.warp_precompute(I0, I1, u1, u2)
    ref = oracle_f64.warp_precompute(I0, I1, u1, u2)
    # validity box identical: exactly zero outside (src/bicubic_interpolation.cpp:214-215)
    uu, vv = xx + u1.astype(np.float64), yy + u2.astype(np.float64)
    inside = (uu >= 1) & (uu < nx - 2) & (vv >= 1) & (vv < ny - 2)
    assert np.all(got["I1wx"][~inside] == 0) and np.all(got["I1wy"][~inside] == 0)
    assert np.all(got["grad"][~inside] == 0)
    assert np.array_equal(got["rho_c"][~inside], -I0[~inside])
    # inside: bicubic of values ~255 with fp32 weights; rho_c multiplies by |u| <= 6
    assert np.abs(got["I1wx"] - ref["I1wx"]).max() < 2e-4
    assert np.abs(got["I1wy"] - ref["I1wy"]).max() < 2e-4
    assert np.abs(got["rho_c"] - ref["rho_c"]).max() < 2e-3
    assert np.allclose(got["grad"], ref["grad"], rtol=1e-4, atol=1e-4)


def test_warp_precompute_large_flow_takes_the_global_gather_path(gpu, oracle_f64):
    """Flows beyond the +-8 px box staged in shared memory (and a flow discontinuity inside one tile)
    must give the same numbers through the out-of-line global gather."""
    ny, nx = 96, 200
    rs = np.random.RandomState(16)
    yy, xx = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    I1 = (128 + 60 * np.sin(0.11 * xx + 0.07 * yy) + 40 * np.cos(0.09 * yy - 0.03 * xx)).astype(np.float32)
    I0 = (I1 + rs.uniform(-5, 5, I1.shape)).astype(np.float32)
    u1 = rs.uniform(-40, 40, I1.shape).astype(np.float32)
    u2 = rs.uniform(-25, 25, I1.shape).astype(np.float32)
    u1[:, :100] = 3.25          # half of every tile row inside the margin, half far outside
    got = gpu.warp_precompute(I0, I1, u1, u2)
    ref = oracle_f64.warp_precompute(I0, I1, u1, u2)
    assert np.abs(got["I1wx"] - ref["I1wx"]).max() < 2e-4
    assert np.abs(got["I1wy"] - ref["I1wy"]).max() < 2e-4
    assert np.abs(got["rho_c"] - ref["rho_c"]).max() < 2e-2      # |u| up to 40 multiplies the gradient error
    assert np.array_equal(got["grad"] == 0, ref["grad"] == 0)


@pytest.mark.parametrize("shape", [_cases.FUNC_SHAPE, (40, 128), (33, 250), (8, 124), (9, 125)])
@pytest.mark.parametrize("iters", [1, 3])
def test_iterate(gpu, oracle_f64, shape, iters):
    """The fused iteration kernel against src/tvl1flow.cpp:114-181 on arbitrary state (random dual
    variables exercise the boundary rules of divergence / forward_gradient)."""
    rs = np.random.RandomState(8)
    f = lambda lo, hi: rs.uniform(lo, hi, shape).astype(np.float32)
    u1, u2 = f(-3, 3), f(-3, 3)
    p = [f(-1, 1) for _ in range(4)]
    ix, iy = f(-20, 20), f(-20, 20)
    ix[rs.uniform(size=shape) < 0.1] = 0.0
    iy[ix == 0] = 0.0                      # exercises grad < GRAD_IS_ZERO
    grad = (ix * ix + iy * iy).astype(np.float32)
    rho_c = f(-30, 30)
    args = (u1, u2, *p, rho_c, ix, iy, grad, 0.25, 0.15, 0.3, iters)
    got = gpu.iterate(*args)
    ref = oracle_f64.iterate(*args)
    for k, name in enumerate(("u1", "u2", "p11", "p12", "p21", "p22")):
        assert np.abs(got[k] - ref[k]).max() < 2e-4 * iters, name
    assert np.allclose(got[6], ref[6], rtol=1e-4)


def test_iterate_keeps_dual_boundary_invariants(gpu):
    """forward_gradient is 0 on the last column / row, so p11,p21 stay 0 there when they start at 0
    (the reference relies on this when it drops those terms in divergence)."""
    shape = (21, 37)
    rs = np.random.RandomState(9)
    f = lambda lo, hi: rs.uniform(lo, hi, shape).astype(np.float32)
    z = np.zeros(shape, np.float32)
    ix, iy = f(-20, 20), f(-20, 20)
    out = gpu.iterate(f(-1, 1), f(-1, 1), z, z, z, z, f(-30, 30), ix, iy, ix * ix + iy * iy,
                      0.25, 0.15, 0.3, 4)
    assert np.all(out[2][:, -1] == 0) and np.all(out[4][:, -1] == 0)
    assert np.all(out[3][-1, :] == 0) and np.all(out[5][-1, :] == 0)


def _iterate_inputs(shape, seed=8):
    rs = np.random.RandomState(seed)
    f = lambda lo, hi: rs.uniform(lo, hi, shape).astype(np.float32)
    u1, u2 = f(-3, 3), f(-3, 3)
    p = [f(-1, 1) for _ in range(4)]
    ix, iy = f(-20, 20), f(-20, 20)
    ix[rs.uniform(size=shape) < 0.1] = 0.0
    iy[ix == 0] = 0.0
    grad = (ix * ix + iy * iy).astype(np.float32)
    rho_c = f(-30, 30)
    return u1, u2, p, rho_c, ix, iy, grad


@pytest.mark.parametrize("shape,cluster", [
    ((68, 120), 0), ((68, 120), 1), ((68, 120), 2), ((30, 40), 1), ((30, 40), 4), ((47, 61), 2),
    ((135, 240), 0), ((135, 240), 4), ((135, 240), 8), ((144, 240), 16), ((270, 480), 0), ((109, 256), 0),
    ((32, 23), 16),
])
def test_iterate_resident_fixed_count(gpu, oracle_f64, shape, cluster):
    """The cluster-resident kernel (whole loop on chip, DSMEM halos) against src/tvl1flow.cpp:114-181
    for a fixed number of passes, for every cluster size."""
    u1, u2, p, rho_c, ix, iy, grad = _iterate_inputs(shape)
    iters = 5
    got = gpu.iterate_resident(u1, u2, *p, rho_c, ix, iy, 0.25, 0.15, 0.3, -1.0, iters, cluster)
    ref = oracle_f64.iterate(u1, u2, *p, rho_c, ix, iy, grad, 0.25, 0.15, 0.3, iters)
    assert got[6] == iters
    if cluster:
        assert got[8] == cluster
    for k, name in enumerate(("u1", "u2", "p11", "p12", "p21", "p22")):
        assert np.abs(got[k] - ref[k]).max() < 2e-4 * iters, (name, got[8])
    assert np.allclose(got[7], ref[6], rtol=1e-4)


def test_iterate_resident_matches_streaming_kernel(gpu):
    u1, u2, p, rho_c, ix, iy, grad = _iterate_inputs((68, 120), seed=12)
    a = gpu.iterate_resident(u1, u2, *p, rho_c, ix, iy, 0.25, 0.15, 0.3, -1.0, 4, 1)
    b = gpu.iterate(u1, u2, *p, rho_c, ix, iy, grad, 0.25, 0.15, 0.3, 4)
    for k in range(6):
        assert np.abs(a[k] - b[k]).max() < 1e-5
    assert np.allclose(a[7], b[6], rtol=1e-6)


@pytest.mark.parametrize("cluster", [1, 2, 4])
def test_iterate_resident_stopping_rule(gpu, oracle_f64, cluster):
    """Stops after the first pass whose mean squared update is <= eps^2 (src/tvl1flow.cpp:113),
    decided on chip, identically for every cluster size."""
    I0, I1 = _cases.synth.make_pair(120, 68, seed=5, scale=0.2)
    z = np.zeros_like(I0)
    c = oracle_f64.warp_precompute(I0, I1, z, z)
    eps = 0.05
    ref = oracle_f64.iterate(z, z, z, z, z, z, c["rho_c"], c["I1wx"], c["I1wy"], c["grad"], 0.25, 0.15, 0.3, 300)
    n_ref = int(np.argmax(ref[6] <= eps * eps)) + 1 if np.any(ref[6] <= eps * eps) else 300
    got = gpu.iterate_resident(z, z, z, z, z, z, c["rho_c"], c["I1wx"], c["I1wy"], 0.25, 0.15, 0.3, eps, 300, cluster)
    assert 1 < n_ref < 300
    assert got[6] == n_ref, (got[6], n_ref)
    assert np.allclose(got[7], ref[6][:n_ref], rtol=1e-3)
    # cap
    got = gpu.iterate_resident(z, z, z, z, z, z, c["rho_c"], c["I1wx"], c["I1wy"], 0.25, 0.15, 0.3, 0.0, 37, cluster)
    assert got[6] == 37


@pytest.mark.parametrize("shape", [(96, 128), (75, 131), (270, 480), (33, 70)])
@pytest.mark.parametrize("iters", [4, 10, 13])
def test_iterate_temporally_blocked_fixed_count(gpu, oracle_f64, shape, iters):
    """k_iterate_tb (2-D TMA halo tiles, up to 4 iterations per launch in shared memory) against
    src/tvl1flow.cpp:114-181 for a fixed number of passes; blocks of 4 plus a shorter last block."""
    u1, u2, p, rho_c, ix, iy, grad = _iterate_inputs(shape, seed=21)
    got = gpu.iterate_loop(u1, u2, *p, rho_c, ix, iy, 0.25, 0.15, 0.3, -1.0, iters, temporal_blocking=2)
    ref = oracle_f64.iterate(u1, u2, *p, rho_c, ix, iy, grad, 0.25, 0.15, 0.3, iters)
    assert got[6] == iters
    assert got[8] == -(-iters // 4), "launches: %d" % got[8]          # ceil(iters / 4) blocks
    for k, name in enumerate(("u1", "u2", "p11", "p12", "p21", "p22")):
        assert np.abs(got[k] - ref[k]).max() < 2e-4 * iters, name
    assert np.isclose(got[7], ref[6][-1], rtol=1e-4)


@pytest.mark.parametrize("shape", [(96, 128), (75, 131), (270, 480), (33, 70), (17, 119), (130, 121), (64, 241), (5, 9), (1, 300), (300, 1)])
@pytest.mark.parametrize("iters", [2, 7, 12])
def test_iterate_two_per_launch_fixed_count(gpu, oracle_f64, shape, iters):
    """k_iterate_t2 (two iterations per launch in registers: the marching kernel with a second iteration one row
    behind the first, 120 owned columns + a halo lane either side per warp) for a fixed number of passes: the same
    bits as one iteration per launch (k_iterate_t1), the reference's values within fp32, half the launches.  Shapes
    around the strip width (119 / 120 / 121 / 241 columns), strips that end inside the image, degenerate images."""
    u1, u2, p, rho_c, ix, iy, grad = _iterate_inputs(shape, seed=23)
    got = gpu.iterate_loop(u1, u2, *p, rho_c, ix, iy, 0.25, 0.15, 0.3, -1.0, iters, temporal_blocking=4)
    one = gpu.iterate_loop(u1, u2, *p, rho_c, ix, iy, 0.25, 0.15, 0.3, -1.0, iters, temporal_blocking=0)
    ref = oracle_f64.iterate(u1, u2, *p, rho_c, ix, iy, grad, 0.25, 0.15, 0.3, iters)
    assert got[6] == iters and one[6] == iters
    assert got[8] == -(-iters // 2), "launches: %d" % got[8]          # ceil(iters / 2) blocks
    for k, name in enumerate(("u1", "u2", "p11", "p12", "p21", "p22")):
        assert np.array_equal(got[k], one[k]), name
        assert np.abs(got[k] - ref[k]).max() < 2e-4 * iters, name
    assert np.isclose(got[7], ref[6][-1], rtol=1e-4)


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("eps", [0.05, 0.02])
def test_iterate_loop_exact_stop(gpu, oracle_f64, mode, eps):
    """Exact stopping under temporal blocking: whatever mix of single iterations, blocks and replays
    the device chooses, the loop must end after the same iteration as the reference and leave the
    same state (mode 2 starts with a full block of 4, which forces the overshoot / replay path when
    the loop is short)."""
    I0, I1 = _cases.synth.make_pair(200, 136, seed=5, scale=0.2)
    z = np.zeros_like(I0)
    c = oracle_f64.warp_precompute(I0, I1, z, z)
    ref = oracle_f64.iterate(z, z, z, z, z, z, c["rho_c"], c["I1wx"], c["I1wy"], c["grad"], 0.25, 0.15, 0.3, 300)
    n_ref = int(np.argmax(ref[6] <= eps * eps)) + 1 if np.any(ref[6] <= eps * eps) else 300
    exact = oracle_f64.iterate(z, z, z, z, z, z, c["rho_c"], c["I1wx"], c["I1wy"], c["grad"], 0.25, 0.15, 0.3, n_ref)
    got = gpu.iterate_loop(z, z, z, z, z, z, c["rho_c"], c["I1wx"], c["I1wy"], 0.25, 0.15, 0.3, eps, 300, mode)
    assert 1 < n_ref < 300
    assert got[6] == n_ref, (got[6], n_ref, got[8])
    for k in range(6):
        assert np.abs(got[k] - exact[k]).max() < 2e-4 * n_ref
    assert np.isclose(got[7], ref[6][n_ref - 1], rtol=1e-3)
    if mode == 0:
        assert got[8] == n_ref
    else:
        assert got[8] <= n_ref


def test_kernel_variants_agree_bitwise(oracle_f64):
    """Which iteration kernel serves a level depends on level and batch size (streaming, temporally
    blocked, cluster-resident); the flow must not: same bits from every combination."""
    import os
    I0, I1 = _cases.synth.make_pair(352, 264, seed=17, scale=0.6)
    kw = dict(nscales=3, warps=3, eps=0.01)
    saved = {k: os.environ.pop(k, None) for k in ("TVL1_NO_TB", "TVL1_NO_RESIDENT", "TVL1_T2")}
    results = []
    try:
        for env in ({}, {"TVL1_NO_TB": "1"}, {"TVL1_NO_RESIDENT": "1"}, {"TVL1_NO_TB": "1", "TVL1_NO_RESIDENT": "1"},
                    {"TVL1_NO_TB": "1", "TVL1_T2": "0"}, {"TVL1_NO_TB": "1", "TVL1_NO_RESIDENT": "1", "TVL1_T2": "0"}):
            for k in ("TVL1_NO_TB", "TVL1_NO_RESIDENT", "TVL1_T2"):
                os.environ.pop(k, None)
            os.environ.update(env)
            g = pkg.TVL1(device=0)           # the switches are read when the context is made
            results.append(g.Dual_TVL1_optic_flow_multiscale(I0, I1, **kw))
            g.close()
    finally:
        for k, v in saved.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v
    a = results[0]
    for r in results[1:]:
        assert np.array_equal(a[2], r[2])
        assert np.array_equal(a[0], r[0]) and np.array_equal(a[1], r[1])


@pytest.mark.parametrize("nx,ny,B", [(250, 190, 5), (121, 67, 9)])
def test_two_per_launch_kernel_in_the_solver(nx, ny, B, monkeypatch):
    """The solver with its streamed levels on the two-iterations-per-launch kernel (TVL1_T2=2 forces it where the
    launch would be too small to choose it) against one iteration per launch: the device's block predictions, the
    rejected blocks and their replays must leave every pair with the same iteration counts and the same bits."""
    pairs = [_cases.synth.make_pair(nx, ny, seed=40 + b, scale=0.3 + 0.1 * (b % 3)) for b in range(B)]
    I0 = np.stack([p[0] for p in pairs])
    I1 = np.stack([p[1] for p in pairs])
    # a pair without motion stops after ONE iteration of every loop: the two-iteration first block of each level (duals
    # taken as zero) overshoots, is rejected, and its first iteration is replayed from zero duals again
    I1[1] = I0[1]
    kw = dict(nscales=3, warps=3, eps=0.01)
    monkeypatch.setenv("TVL1_NO_RESIDENT", "1")
    monkeypatch.setenv("TVL1_NO_TB", "1")
    out = []
    for t2 in ("0", "2"):
        monkeypatch.setenv("TVL1_T2", t2)
        g = pkg.TVL1(device=0, max_batch=B)
        out.append(g.Dual_TVL1_optic_flow_multiscale(I0, I1, **kw))
        st = g.stats()
        out[-1] = out[-1] + (sum(st["level_iterate_launches"]),)
        g.close()
    a, b = out
    assert np.array_equal(a[2], b[2]), (a[2].tolist(), b[2].tolist())
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.all(a[2][1] == 1)             # the still pair: one iteration per loop
    assert b[4] != a[4]                     # the blocked path really ran (launch counts differ)


def test_iterate_loop_variants_agree_bitwise(gpu, oracle_f64):
    """Fixed 13 iterations on one level: single iterations (mode 0) against blocks of 4 (mode 2)."""
    I0, I1 = _cases.synth.make_pair(200, 136, seed=5, scale=0.2)
    z = np.zeros_like(I0)
    c = oracle_f64.warp_precompute(I0, I1, z, z)
    a = gpu.iterate_loop(z, z, z, z, z, z, c["rho_c"], c["I1wx"], c["I1wy"], 0.25, 0.15, 0.3, -1.0, 13, 0)
    b = gpu.iterate_loop(z, z, z, z, z, z, c["rho_c"], c["I1wx"], c["I1wy"], 0.25, 0.15, 0.3, -1.0, 13, 2)
    t2 = gpu.iterate_loop(z, z, z, z, z, z, c["rho_c"], c["I1wx"], c["I1wy"], 0.25, 0.15, 0.3, -1.0, 13, 4)
    assert a[6] == b[6] == t2[6] == 13
    for k in range(6):
        assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a[k], t2[k]), k


def test_iterate_loop_replay_on_immediate_stop(gpu, oracle_f64):
    """A loop that stops after its first iteration although a block of 4 was started."""
    shape = (64, 96)
    z = np.zeros(shape, np.float32)
    got = gpu.iterate_loop(z, z, z, z, z, z, z, z, z, 0.25, 0.15, 0.3, 0.01, 300, 2)   # zero update -> error 0
    assert got[6] == 1 and got[7] == 0.0 and got[8] == 2       # block of 4 rejected, replay of 1 accepted


# ---- the solver against the golden vectors of the unmodified reference --------------------------

@pytest.mark.parametrize("name", sorted(_cases.SOLVER_CASES))
def test_solver_vs_golden(gpu, golden, name):
    case = _cases.SOLVER_CASES[name]
    I0, I1 = _cases.solver_inputs(case)
    u1, u2, iters, errs = gpu.Dual_TVL1_optic_flow_multiscale(I0, I1, **case["kw"])
    pre = "f64/%s/" % name
    assert np.array_equal(iters, golden[pre + "iters"]), (iters.tolist(), golden[pre + "iters"].tolist())
    assert_flow_close(u1, u2, golden[pre + "u1"], golden[pre + "u2"], name)
    assert np.allclose(errs, golden[pre + "errs"], rtol=1e-3, atol=1e-6)


def test_solver_f64_entry_matches_f32_entry(gpu):
    case = _cases.SOLVER_CASES["ms_64x48"]
    I0, I1 = _cases.solver_inputs(case)
    a = gpu.Dual_TVL1_optic_flow_multiscale(I0, I1, **case["kw"])
    b = gpu.Dual_TVL1_optic_flow_multiscale(I0.astype(np.float64), I1.astype(np.float64), **case["kw"])
    assert b[0].dtype == np.float64
    assert np.array_equal(a[0].astype(np.float64), b[0]) and np.array_equal(a[2], b[2])


def test_single_scale_entry(gpu, oracle_f64):
    """Dual_TVL1_optic_flow: the initial flow is used (src/tvl1flow.cpp:94)."""
    I0, I1 = _cases.synth.make_pair(96, 72, seed=21, scale=0.3)
    rs = np.random.RandomState(2)
    u1 = rs.uniform(-0.5, 0.5, I0.shape).astype(np.float32)
    u2 = rs.uniform(-0.5, 0.5, I0.shape).astype(np.float32)
    g = gpu.Dual_TVL1_optic_flow(I0, I1, u1, u2, warps=3, eps=0.01)
    r = oracle_f64.single_scale(I0, I1, u1, u2, warps=3, eps=0.01)
    assert np.array_equal(g[2], r[2]), (g[2], r[2])
    assert_flow_close(g[0], g[1], r[0], r[1], "single scale")


def test_large_motion_solve(gpu, oracle_f64):
    """A pair whose motion (about 12 px at the finest level) exceeds the warp kernel's staged margin:
    the solver must still agree with the reference."""
    I0, I1 = _cases.synth.make_pair(320, 240, seed=9, scale=4.0)
    kw = dict(nscales=4, warps=4, eps=0.01)
    u1, u2, iters, _ = gpu.Dual_TVL1_optic_flow_multiscale(I0, I1, **kw)
    r1, r2, riters, _ = oracle_f64.multiscale(I0, I1, **kw)
    assert np.abs(r1).max() > 8.5          # the case really leaves the margin
    assert np.array_equal(iters, riters), (iters.tolist(), riters.tolist())
    assert_flow_close(u1, u2, r1, r2, "large motion")


def test_tiny_and_thin_images(gpu, oracle_f64):
    """Levels narrower than one warp strip / one warp tile, heights below one strip."""
    for (nx, ny, ns) in [(24, 40, 2), (125, 9, 1), (16, 16, 1), (300, 20, 2)]:
        I0, I1 = _cases.synth.make_pair(nx, ny, seed=3, scale=0.2)
        kw = dict(nscales=ns, warps=2, eps=0.01)
        u1, u2, iters, _ = gpu.Dual_TVL1_optic_flow_multiscale(I0, I1, **kw)
        r1, r2, riters, _ = oracle_f64.multiscale(I0, I1, **kw)
        assert np.array_equal(iters, riters), (nx, ny, iters.tolist(), riters.tolist())
        assert_flow_close(u1, u2, r1, r2, "%dx%d" % (nx, ny))


def _sweep_cases():
    rs = np.random.RandomState(2026)
    cases = []
    for k in range(14):
        nx, ny = int(rs.randint(48, 330)), int(rs.randint(40, 260))
        zf = float(rs.choice([0.5, 0.5, 0.6, 0.7, 0.4, 0.8]))
        ns = int(rs.randint(1, 5))
        # the CLI's clamp keeps the coarsest level >= ~16 px (tvl1flow_main.cpp:185-188)
        ns = min(ns, pkg.clamp_nscales(nx, ny, ns, zf))
        cases.append(dict(nx=nx, ny=ny, seed=int(rs.randint(1, 10000)), scale=float(rs.uniform(0.1, 0.8)),
                          kw=dict(tau=float(rs.choice([0.25, 0.2, 0.1])), lam=float(rs.choice([0.15, 0.05, 0.3])),
                                  theta=float(rs.choice([0.3, 0.2, 0.5])), nscales=ns, zfactor=zf,
                                  warps=int(rs.randint(1, 5)), eps=float(rs.choice([0.01, 0.02, 0.005])))))
    return cases


@pytest.mark.parametrize("case", _sweep_cases(), ids=lambda c: "%dx%d_z%.1f_s%d_w%d" % (
    c["nx"], c["ny"], c["kw"]["zfactor"], c["kw"]["nscales"], c["kw"]["warps"]))
def test_random_parameter_sweep(gpu, oracle_f64, case):
    """Seeded sweep over sizes (odd widths, non-multiples of 4), pyramid factors (general bicubic
    zoom_out path, wider Gaussian windows), step sizes and thresholds: iteration counts and flow
    against the oracle."""
    I0, I1 = _cases.synth.make_pair(case["nx"], case["ny"], seed=case["seed"], scale=case["scale"])
    u1, u2, iters, _ = gpu.Dual_TVL1_optic_flow_multiscale(I0, I1, **case["kw"])
    r1, r2, riters, _ = oracle_f64.multiscale(I0, I1, **case["kw"])
    assert np.array_equal(iters, riters), (iters.tolist(), riters.tolist())
    assert_flow_close(u1, u2, r1, r2, str(case))


def test_eps_zero_runs_to_the_cap(gpu):
    I0, I1 = _cases.synth.make_pair(48, 40, seed=3, scale=0.3)
    _, _, iters, _ = gpu.Dual_TVL1_optic_flow_multiscale(I0, I1, nscales=2, warps=2, eps=0.0)
    assert np.all(iters == 300)     # MAX_ITERATIONS, src/tvl1flow.cpp:22,113


def test_sigma_too_large_is_an_error(gpu):
    I0, I1 = _cases.synth.make_pair(40, 32, seed=3)
    with pytest.raises(pkg.TVL1Error) as e:
        gpu.Dual_TVL1_optic_flow_multiscale(I0, I1, nscales=5)   # level 3 is 5 px wide < 6 taps
    assert e.value.code == 2


def test_identical_images_give_zero_flow(gpu):
    I0, _ = _cases.synth.make_pair(128, 96, seed=8)
    u1, u2, iters, errs = gpu.Dual_TVL1_optic_flow_multiscale(I0, I0, nscales=3)
    assert np.all(u1 == 0) and np.all(u2 == 0)
    assert np.all(iters == 1) and np.all(errs == 0)


def test_batch_equals_individual_solves(gpu):
    """Pairs of a batch are independent units: batching must not change a single bit."""
    pairs = [_cases.synth.make_pair(100, 76, seed=40 + b, scale=0.4) for b in range(5)]
    I0 = np.stack([p[0] for p in pairs])
    I1 = np.stack([p[1] for p in pairs])
    kw = dict(nscales=3, warps=3, eps=0.01)
    bu1, bu2, bit, _ = gpu.Dual_TVL1_optic_flow_multiscale(I0, I1, **kw)
    for b in range(5):
        u1, u2, it, _ = gpu.Dual_TVL1_optic_flow_multiscale(I0[b], I1[b], **kw)
        assert np.array_equal(it, bit[b])
        assert np.array_equal(u1, bu1[b]) and np.array_equal(u2, bu2[b])
    # chunking by max_batch must not matter either
    gpu.set_max_batch(2)
    cu1, cu2, cit, _ = gpu.Dual_TVL1_optic_flow_multiscale(I0, I1, **kw)
    gpu.set_max_batch(32)
    assert np.array_equal(cu1, bu1) and np.array_equal(cu2, bu2) and np.array_equal(cit, bit)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_frame_sequence_equals_pairwise(gpu, dt):
    """Video mode: F frames -> F-1 flows, each frame uploaded once; must equal the per-pair solves
    bit for bit, also when the sequence is cut into chunks (the boundary frame is shared)."""
    F = 6
    base = [_cases.synth.make_pair(97, 61, seed=90 + k, scale=0.3)[0] for k in range(2)]
    frames = np.stack([np.roll(base[k % 2], (k, 2 * k), axis=(0, 1)) for k in range(F)]).astype(dt)
    kw = dict(nscales=3, warps=2, eps=0.01)
    for mb in (32, 2):
        gpu.set_max_batch(mb)
        su1, su2, sit, _ = gpu.solve_sequence(frames, **kw)
        assert su1.dtype == dt and su1.shape == (F - 1, 61, 97)
        for b in range(F - 1):
            u1, u2, it, _ = gpu.Dual_TVL1_optic_flow_multiscale(frames[b], frames[b + 1], **kw)
            assert np.array_equal(it, sit[b])
            assert np.array_equal(u1, su1[b]) and np.array_equal(u2, su2[b])
    gpu.set_max_batch(32)
    with pytest.raises(pkg.TVL1Error):
        gpu.solve_sequence(frames[:1], **kw)


def test_frame_sequence_of_8_bit_frames(gpu):
    """tvl1_solve_sequence_u8: 8-bit frames cross PCIe as bytes and are widened on the device -- the same bits
    as the fp32 sequence call on the widened frames, whole and cut into chunks, at a size whose pixel count is
    not a multiple of four."""
    F = 7
    base = [_cases.synth.make_pair(97, 61, seed=190 + k, scale=0.3)[0] for k in range(2)]
    frames = np.stack([np.roll(base[k % 2], (k, 2 * k), axis=(0, 1)) for k in range(F)])
    q = np.clip(np.rint(frames), 0, 255).astype(np.uint8)
    kw = dict(nscales=3, warps=2, eps=0.01)
    for mb in (32, 2):
        gpu.set_max_batch(mb)
        a = gpu.solve_sequence(q, **kw)
        b = gpu.solve_sequence(q.astype(np.float32), **kw)
        assert a[0].dtype == np.float32 and a[0].shape == (F - 1, 61, 97)
        assert np.array_equal(a[2], b[2]) and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    gpu.set_max_batch(32)
    with pytest.raises(pkg.TVL1Error):
        gpu.solve_sequence(q[:1], **kw)


def test_ragged_batch_keeps_both_workspaces(gpu):
    """A batch that is not a multiple of the lock-step size alternates between two batch sizes; the
    displaced workspace and its solve graph are kept and swapped back in (profiles/run_e2e.py shows
    the time this saves); results must not depend on which of the two a chunk ran in."""
    pairs = [_cases.synth.make_pair(160, 120, seed=300 + b, scale=0.4) for b in range(7)]
    I0 = np.stack([p[0] for p in pairs])
    I1 = np.stack([p[1] for p in pairs])
    kw = dict(nscales=3, warps=2, eps=0.01)
    g = pkg.TVL1(device=0, max_batch=3)          # chunks of 3, 3, 1
    g.set_lanes(host_lanes=1)
    a = g.Dual_TVL1_optic_flow_multiscale(I0, I1, **kw)
    b = g.Dual_TVL1_optic_flow_multiscale(I0, I1, **kw)
    c = g.Dual_TVL1_optic_flow_multiscale(I0[:4], I1[:4], **kw)      # chunks of 3, 1 again
    g.close()
    assert np.array_equal(c[0], a[0][:4]) and np.array_equal(c[2], a[2][:4])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    for k in range(7):
        u1, u2, it, _ = gpu.Dual_TVL1_optic_flow_multiscale(I0[k], I1[k], **kw)
        assert np.array_equal(u1, a[0][k]) and np.array_equal(u2, a[1][k]) and np.array_equal(it, a[2][k])


@pytest.mark.parametrize("npairs,max_batch,lanes,short_div", [(24, 4, 3, 2), (21, 4, 3, 2), (24, 4, 4, 4), (13, 6, 2, 0),
                                                              (40, 4, 8, 2), (9, 2, 4, 2)])
def test_host_batch_chunk_schedules(gpu, npairs, max_batch, lanes, short_div, monkeypatch):
    """Host-buffer batches are cut into lock-step chunks that lanes take from a shared queue -- short first /
    last chunks when that adds no third chunk size, ragged remainders, more lanes than chunks, no temporal
    blocking while lanes share the GPU: whatever the schedule, every pair's flow and iteration counts are the
    bits of its individual solve."""
    monkeypatch.setenv("TVL1_SHORT_DIV", str(short_div))
    pairs = [_cases.synth.make_pair(128, 96, seed=500 + b, scale=0.3 + 0.02 * (b % 5)) for b in range(npairs)]
    I0 = np.stack([p[0] for p in pairs])
    I1 = np.stack([p[1] for p in pairs])
    kw = dict(nscales=3, warps=2, eps=0.01)
    g = pkg.TVL1(device=0, max_batch=max_batch)
    g.set_lanes(host_lanes=lanes)
    a = g.Dual_TVL1_optic_flow_multiscale(I0, I1, **kw)
    b = g.Dual_TVL1_optic_flow_multiscale(I0, I1, **kw)          # second call: workspaces and graphs re-used
    g.close()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    for k in range(0, npairs, 3):
        u1, u2, it, _ = gpu.Dual_TVL1_optic_flow_multiscale(I0[k], I1[k], **kw)
        assert np.array_equal(u1, a[0][k]) and np.array_equal(u2, a[1][k]) and np.array_equal(it, a[2][k]), k


# The call-wide pipeline is opt-in (TVL1_HOST_PIPE=1): an intermittent launch failure was seen with it on the 256 x 1080p
# bench workload, and a launch failure would poison the CUDA context of the whole test process.  Its tests
# (always green so far) therefore run on request only: TVL1_TEST_HOST_PIPE=1 pytest -m gpu -k pinned_host_batch_pipeline
_PIPE_TESTS = pytest.mark.skipif(os.environ.get("TVL1_TEST_HOST_PIPE") != "1",
                                 reason="opt-in host-buffer pipeline: set TVL1_TEST_HOST_PIPE=1")


@_PIPE_TESTS
@pytest.mark.parametrize("npairs,max_batch,lanes,dtype,form", [
    (29, 8, 3, "float32", "pairs"), (29, 8, 3, "float64", "pairs"), (16, 4, 2, "float32", "sequence"),
    (13, 4, 4, "float64", "sequence"), (40, 16, 1, "float32", "pairs"), (9, 2, 8, "float32", "pairs"),
    (14, 4, 3, "uint8", "sequence")])
def test_pinned_host_batch_pipeline(gpu, npairs, max_batch, lanes, dtype, form, monkeypatch):
    """TVL1_HOST_PIPE=1: pinned host buffers go through the call-wide upload -> solve -> download pipeline (solve_host_pipelined):
    ramped chunk sizes (tvl1_plan_chunks), one ordered copy stream each way, a ring of device slots that chunks
    re-use.  Whatever the schedule, fp32 or fp64 buffers, pairs or a frame sequence: every pair's flow and iteration
    counts are the bits of its individual solve, call after call."""
    import torch
    monkeypatch.setenv("TVL1_HOST_PIPE", "1")
    monkeypatch.setenv("TVL1_MIN_CHUNK", "1")                  # ramps down to single pairs: every chunk size the lanes can meet
    nx, ny = 128, 96
    kw = dict(nscales=3, warps=2, eps=0.01)
    tdt = torch.float64 if dtype == "float64" else torch.float32          # 8-bit frames give fp32 flows
    if form == "pairs":
        pairs = [_cases.synth.make_pair(nx, ny, seed=700 + b, scale=0.3 + 0.02 * (b % 5)) for b in range(npairs)]
        A = np.stack([p[0] for p in pairs]).astype(dtype)
        Bm = np.stack([p[1] for p in pairs]).astype(dtype)
        hA = torch.from_numpy(A).pin_memory()
        hB = torch.from_numpy(Bm).pin_memory()
    else:
        fr = [_cases.synth.make_pair(nx, ny, seed=800 + b, scale=0.3)[b & 1] for b in range(npairs + 1)]
        A = np.stack(fr)
        A = np.clip(np.rint(A), 0, 255).astype(np.uint8) if dtype == "uint8" else A.astype(dtype)
        hA = torch.from_numpy(A).pin_memory()
    assert len(pkg.tvl1.plan_chunks(npairs, max_batch)) > 1
    g = pkg.TVL1(device=0, max_batch=max_batch)
    g.set_lanes(host_lanes=lanes)
    outs = []
    for rep in range(2):                                        # second call: slots, workspaces and graphs re-used
        hu1 = torch.full((npairs, ny, nx), float("nan"), dtype=tdt).pin_memory()
        hu2 = torch.full((npairs, ny, nx), float("nan"), dtype=tdt).pin_memory()
        it = np.zeros((npairs, kw["nscales"], kw["warps"]), np.int32)
        er = np.zeros((npairs, kw["nscales"], kw["warps"]), np.float64)
        if form == "pairs":
            g.solve_batch_host_ptr(hA.data_ptr(), hB.data_ptr(), hu1.data_ptr(), hu2.data_ptr(), npairs, nx, ny,
                                   dtype=dtype, iters=it, errs=er, **kw)
        else:
            g.solve_sequence_host_ptr(hA.data_ptr(), hu1.data_ptr(), hu2.data_ptr(), npairs + 1, nx, ny,
                                      dtype=dtype, iters=it, errs=er, **kw)
        outs.append((hu1.numpy().copy(), hu2.numpy().copy(), it, er))
    g.close()
    a, b = outs
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert np.isfinite(a[0]).all() and np.isfinite(a[1]).all()
    for k in range(npairs):
        I0k, I1k = (A[k], Bm[k]) if form == "pairs" else (A[k], A[k + 1])
        if dtype == "uint8":
            I0k, I1k = I0k.astype(np.float32), I1k.astype(np.float32)
        u1, u2, itk, erk = gpu.Dual_TVL1_optic_flow_multiscale(I0k, I1k, **kw)
        assert np.array_equal(u1, a[0][k]) and np.array_equal(u2, a[1][k]), k
        assert np.array_equal(itk, a[2][k]) and np.array_equal(erk, a[3][k]), k


@pytest.mark.parametrize("env", [{}, {"TVL1_NO_RESIDENT": "1"}], ids=["default", "all_levels_streamed"])
def test_pinned_host_batch_streamed_levels_in_small_chunks(gpu, env, monkeypatch):
    """Host-buffer batch of mid-size images (640x360: the finest level streams through HBM, the next ones live on chip;
    with TVL1_NO_RESIDENT=1 every level streams, down to 160x90) cut into chunks of 1..4 pairs on three lanes: small
    streamed levels inside concurrently solved chunks, against the individual solves."""
    import torch
    for k_, v_ in env.items():
        monkeypatch.setenv(k_, v_)
    nx, ny, npairs = 640, 360, 14
    kw = dict(nscales=3, warps=2, eps=0.01)
    pairs = [_cases.synth.make_pair(nx, ny, seed=950 + b, scale=0.5) for b in range(npairs)]
    A = np.stack([p[0] for p in pairs])
    Bm = np.stack([p[1] for p in pairs])
    hA, hB = torch.from_numpy(A).pin_memory(), torch.from_numpy(Bm).pin_memory()
    g = pkg.TVL1(device=0, max_batch=4)
    g.set_lanes(host_lanes=3)
    hu1 = torch.zeros((npairs, ny, nx)).pin_memory()
    hu2 = torch.zeros((npairs, ny, nx)).pin_memory()
    it = np.zeros((npairs, 3, 2), np.int32)
    for rep in range(2):
        g.solve_batch_host_ptr(hA.data_ptr(), hB.data_ptr(), hu1.data_ptr(), hu2.data_ptr(), npairs, nx, ny,
                               dtype="float32", iters=it, **kw)
    g.close()
    for k_ in env:
        monkeypatch.delenv(k_)
    for k in range(0, npairs, 3):
        u1, u2, itk, _ = gpu.Dual_TVL1_optic_flow_multiscale(A[k], Bm[k], **kw)
        assert np.array_equal(u1, hu1[k].numpy()) and np.array_equal(u2, hu2[k].numpy()), k
        assert np.array_equal(itk, it[k]), k


@_PIPE_TESTS
def test_pinned_host_batch_pipeline_matches_lane_copies(monkeypatch):
    """A/B switch: TVL1_HOST_PIPE=1 selects the call-wide pipeline instead of lanes that do their own copies;
    TVL1_CHUNKS overrides its chunk sizes.  Same bits either way."""
    import torch
    nx, ny, npairs = 160, 120, 21
    kw = dict(nscales=3, warps=2, eps=0.01)
    pairs = [_cases.synth.make_pair(nx, ny, seed=900 + b, scale=0.4) for b in range(npairs)]
    hA = torch.from_numpy(np.stack([p[0] for p in pairs])).pin_memory()
    hB = torch.from_numpy(np.stack([p[1] for p in pairs])).pin_memory()
    res = []
    for env in ({"TVL1_HOST_PIPE": "1"}, {"TVL1_HOST_PIPE": "0"}, {"TVL1_HOST_PIPE": "1", "TVL1_CHUNKS": "1,5,3,2"}):
        for k_, v_ in env.items():
            monkeypatch.setenv(k_, v_)
        g = pkg.TVL1(device=0, max_batch=6)
        g.set_lanes(host_lanes=3)
        hu1 = torch.zeros((npairs, ny, nx)).pin_memory()
        hu2 = torch.zeros((npairs, ny, nx)).pin_memory()
        it = np.zeros((npairs, 3, 2), np.int32)
        g.solve_batch_host_ptr(hA.data_ptr(), hB.data_ptr(), hu1.data_ptr(), hu2.data_ptr(), npairs, nx, ny,
                               dtype="float32", iters=it, **kw)
        g.close()
        for k_ in env:
            monkeypatch.delenv(k_)
        res.append((hu1.numpy().copy(), hu2.numpy().copy(), it))
    for r in res[1:]:
        assert all(np.array_equal(x, y) for x, y in zip(res[0], r))


def test_band_code_path_single_rank(gpu, oracle_f64):
    """Row-band mode with one rank (a band = the whole level, NCCL communicator of size 1): the
    row-window kernels, the all-reduced stopping rule and the in-place all-gather must reproduce the
    ordinary solve bit for bit."""
    I0, I1 = _cases.synth.make_pair(200, 144, seed=31, scale=0.5)
    kw = dict(nscales=3, warps=3, eps=0.01)
    a = gpu.Dual_TVL1_optic_flow_multiscale(I0, I1, **kw)
    gpu.band_init(0, 1, gpu.band_unique_id())
    b = gpu.band_solve(I0, I1, min_split_rows=-60, **kw)     # splits the 144- and 72-row levels
    assert np.array_equal(a[2], b[2]), (a[2].tolist(), b[2].tolist())
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    r = oracle_f64.multiscale(I0, I1, **kw)
    assert np.array_equal(b[2], r[2])
    assert_flow_close(b[0], b[1], r[0], r[1], "band mode")


def test_band_code_path_single_rank_temporal_blocking(gpu):
    """Same with levels large enough for the temporally blocked kernel inside the band (tile grid
    anchored at the band, T-row halo logic, error sums through the mailbox), at a tight epsilon so
    that blocks of 4 iterations, predicted short blocks and exact replays all occur; also the NCCL
    exchange mode on the same context."""
    I0, I1 = _cases.synth.make_pair(320, 264, seed=32, scale=0.5)
    kw = dict(nscales=3, warps=3, eps=0.002)
    a = gpu.Dual_TVL1_optic_flow_multiscale(I0, I1, **kw)
    gpu.band_init(0, 1, gpu.band_unique_id())
    assert gpu.band_exchange_mode() == "peer"
    for mode in ("peer", "nccl", "peer"):
        gpu.band_set_exchange(mode)
        assert gpu.band_exchange_mode() == mode
        b = gpu.band_solve(I0, I1, min_split_rows=-100, **kw)    # splits the 264- and 132-row levels
        assert np.array_equal(a[2], b[2]), (mode, a[2].tolist(), b[2].tolist())
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), mode
    assert gpu.stats()["host_syncs"] == 0          # peer mode: the whole solve without a host round trip
    assert a[2].max() > 8                           # long enough loops for multi-iteration blocks


def _run_band_check(world, *args, timeout=600):
    import socket
    import subprocess
    import sys
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "band_check.py")] + [str(a) for a in args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and line, (r.returncode, r.stdout[-2000:], r.stderr[-3000:])
    import json
    return json.loads(line[-1])


@pytest.mark.parametrize("case", ["small", "4k"])
def test_row_bands_over_all_gpus_of_the_box(case):
    """SURVEY 8e row 2 / BASELINE configs[3]: one image pair split into row bands over
    min(device_count, 8) GPUs, one process per GPU under torch.distributed.run.  Every rank must get
    the single-GPU flow bit for bit with identical iteration counts, in both exchange modes (peer
    memory fused into the iteration kernels; NCCL send/recv + all-reduce per iteration).  Skipped on a
    one-GPU box (the single-rank tests above cover the code path there)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(n, 8)
    if case == "small":
        res = _run_band_check(world, 1024, 768, 4, 3, 0.01, 96)
    else:
        res = _run_band_check(world, 3840, 2160, 6, 10, 0.001, 512)
    assert res["all_ranks_ok"] and res["same_iteration_counts"] and res["both_exchange_modes_match_single_gpu"], res
    assert res["max_abs_flow_diff"] == 0.0, res


def test_dropin_symbols_from_threads(gpu):
    """The reference's mangled C++ entry point (src/tvl1flow.h:56-70, ofpix_t = double) called from
    two host threads at once, one context per thread (and one GPU per thread when there are two)."""
    import ctypes as C
    import threading
    import torch
    lib = C.CDLL(pkg.library_path())
    fn = getattr(lib, "_Z31Dual_TVL1_optic_flow_multiscalePdS_S_S_iidddididb")
    fn.restype = None
    ndev = torch.cuda.device_count()
    pairs = [_cases.synth.make_pair(120, 88, seed=60 + t, scale=0.4) for t in range(2)]
    out, errors = [None, None], []

    def work(t):
        try:
            lib.tvl1_dropin_set_device(C.c_int(t % ndev))
            I0 = pairs[t][0].astype(np.float64)
            I1 = pairs[t][1].astype(np.float64)
            u = np.empty((2,) + I0.shape, np.float64)
            p = lambda a: a.ctypes.data_as(C.c_void_p)
            for _ in range(3):
                fn(p(I0), p(I1), p(u[0]), C.c_void_p(u[1].ctypes.data), C.c_int(120), C.c_int(88), C.c_double(0.25),
                   C.c_double(0.15), C.c_double(0.3), C.c_int(3), C.c_double(0.5), C.c_int(3), C.c_double(0.01),
                   C.c_bool(False))
            out[t] = u
        except Exception as e:          # noqa: BLE001
            errors.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errors, errors
    for t in range(2):
        u1, u2, _, _ = gpu.Dual_TVL1_optic_flow_multiscale(pairs[t][0].astype(np.float64), pairs[t][1].astype(np.float64),
                                                           nscales=3, warps=3, eps=0.01)
        assert np.array_equal(out[t][0], u1) and np.array_equal(out[t][1], u2)


def test_upstream_c99_entry_points(gpu):
    """The C-linkage, all-float names of the upstream library (3rdparty/tvl1flow_3/tvl1flow_lib.c:45-59,
    :299-314) served by the same CUDA path: bit-identical to tvl1_solve_f32 / tvl1_single_scale_f32,
    and within the flow tolerance of the upstream CPU build when that travelled with the snapshot."""
    import ctypes as C
    from oracle import loader
    lib = loader.c99_signature(C.CDLL(pkg.library_path()))
    I0, I1 = _cases.synth.make_pair(176, 132, seed=77, scale=0.6)
    kw = dict(nscales=4, warps=4, eps=0.01)
    u1, u2 = loader.c99_multiscale(lib, I0, I1, **kw)
    g1, g2, _, _ = gpu.Dual_TVL1_optic_flow_multiscale(I0.astype(np.float32), I1.astype(np.float32), **kw)
    assert np.array_equal(u1, g1) and np.array_equal(u2, g2)
    if loader.upstream_c99_available():
        up = loader.c99_signature(C.CDLL(loader.UPSTREAM_C99))
        r1, r2 = loader.c99_multiscale(up, I0, I1, **kw)
        assert_flow_close(u1, u2, r1, r2, "upstream C99 build")
    # one level: u1, u2 are the initial flow on entry
    a = I0.astype(np.float32).copy()
    b = I1.astype(np.float32).copy()
    s = np.full((2,) + a.shape, 0.25, np.float32)
    lib.Dual_TVL1_optic_flow(a.ctypes.data, b.ctypes.data, s[0].ctypes.data, s[1].ctypes.data,
                             176, 132, 0.25, 0.15, 0.3, 3, 0.01, False)
    h1, h2, _, _ = gpu.Dual_TVL1_optic_flow(a, b, np.full(a.shape, 0.25, np.float32),
                                            np.full(a.shape, 0.25, np.float32), warps=3, eps=0.01)
    assert np.array_equal(s[0], h1) and np.array_equal(s[1], h2)


# ---- BASELINE.json configs --------------------------------------------------------------------

def reference_cpu():
    """The compiled reference when it travelled with the snapshot, else the oracle port."""
    kind = "reference" if available("reference", np.float64) else "port"
    return CpuTvl1(kind, np.float64), kind


@pytest.mark.parametrize("nx,ny", [(640, 480), (1024, 436)])
def test_baseline_configs_vs_reference(gpu, nx, ny):
    """configs[0] and configs[1]: default parameters, 5 scales x 5 warps."""
    cpu, kind = reference_cpu()
    I0, I1 = _cases.synth.make_pair(nx, ny, seed=1234)
    kw = dict(tau=0.25, lam=0.15, theta=0.3, nscales=5, zfactor=0.5, warps=5, eps=0.01)
    u1, u2, iters, errs = gpu.Dual_TVL1_optic_flow_multiscale(I0, I1, **kw)
    r1, r2, riters, rerrs = cpu.multiscale(I0, I1, **kw)
    mean, mx = flow_diff(u1, u2, r1, r2)
    print("%dx%d vs %s: mean|d|=%.3g max|d|=%.3g iters=%s" % (nx, ny, kind, mean, mx, iters.tolist()))
    assert np.array_equal(iters, riters), (iters.tolist(), riters.tolist())
    assert mean <= MEAN_TOL and mx <= MAX_TOL


def test_1080p_properties(gpu):
    """Full-size (1920x1080, 5 scales x 5 warps) checks that do not need the CPU oracle:
    run-to-run determinism, batch-position independence, and the zero-motion fixed point."""
    nx, ny = 1920, 1080
    a = _cases.synth.make_pair(nx, ny, seed=1234)
    b = _cases.synth.make_pair(nx, ny, seed=1235)
    I0 = np.stack([a[0], b[0], a[0]])
    I1 = np.stack([a[1], b[1], a[0]])
    u1, u2, iters, errs = gpu.Dual_TVL1_optic_flow_multiscale(I0, I1)
    v1, v2, jters, _ = gpu.Dual_TVL1_optic_flow_multiscale(I0[::-1].copy(), I1[::-1].copy())
    assert np.array_equal(u1, v1[::-1]) and np.array_equal(u2, v2[::-1]) and np.array_equal(iters, jters[::-1])
    assert np.all(u1[2] == 0) and np.all(u2[2] == 0) and np.all(iters[2] == 1)
    assert iters.min() >= 1 and iters.max() <= 300
    # the synthetic motion is (2.3, 1.4) outside the disc and (-3.3, 2.6) inside it
    yy, xx = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    r2 = (xx - 0.5 * nx) ** 2 + (yy - 0.5 * ny) ** 2
    bg = r2 > (0.25 * ny + 40) ** 2
    bg[:40] = bg[-40:] = False
    bg[:, :40] = bg[:, -40:] = False
    disc = r2 < (0.25 * ny - 40) ** 2
    assert abs(np.median(u1[0][bg]) - 2.3) < 0.1 and abs(np.median(u2[0][bg]) - 1.4) < 0.1
    assert abs(np.median(u1[0][disc]) + 3.3) < 0.1 and abs(np.median(u2[0][disc]) - 2.6) < 0.1


def test_1080p_vs_reference(gpu):
    """configs[2] unit of work (one 1080p pair) against the CPU reference / oracle."""
    cpu, kind = reference_cpu()
    I0, I1 = _cases.synth.make_pair(1920, 1080, seed=1234)
    u1, u2, iters, _ = gpu.Dual_TVL1_optic_flow_multiscale(I0, I1)
    r1, r2, riters, _ = cpu.multiscale(I0, I1)
    mean, mx = flow_diff(u1, u2, r1, r2)
    print("1080p vs %s: mean|d|=%.3g max|d|=%.3g iters=%s" % (kind, mean, mx, iters.tolist()))
    assert np.array_equal(iters, riters), (iters.tolist(), riters.tolist())
    assert mean <= MEAN_TOL and mx <= MAX_TOL


def test_8k_vs_reference(gpu):
    """configs[4]: 7680x4320, default parameters (about 7 s of CPU reference on 16 threads)."""
    cpu, kind = reference_cpu()
    cpu.set_threads(os.cpu_count() or 1)
    I0, I1 = _cases.synth.make_pair(7680, 4320, seed=1234)
    u1, u2, iters, _ = gpu.Dual_TVL1_optic_flow_multiscale(I0, I1)
    r1, r2, riters, _ = cpu.multiscale(I0, I1)
    mean, mx = flow_diff(u1, u2, r1, r2)
    print("8K vs %s: mean|d|=%.3g max|d|=%.3g iters=%s" % (kind, mean, mx, iters.tolist()))
    assert np.array_equal(iters, riters), (iters.tolist(), riters.tolist())
    assert mean <= MEAN_TOL and mx <= MAX_TOL


def test_4k_vs_reference(gpu):
    """configs[3]: 3840x2160, 6 scales x 10 warps, eps 0.001 (60 warp steps of up to 300 iterations;
    about 12 s of CPU reference).  Iteration counts, the mean bound and the max bound all hold: measured
    max 7.4e-3 px at one pixel on the edge of the moving disc (row 712, col 1520), where the problem is
    ill-conditioned in fp32 -- the reference's OWN float build differs from its fp64 build by 9.4e-2 px
    there and exceeds 1e-2 at 33 pixels (profiles/diff_4k.py, profiles/r2a_diff_4k.txt); when that build
    travelled with the snapshot the test also checks that the CUDA path is the closer one."""
    cpu, kind = reference_cpu()
    cpu.set_threads(os.cpu_count() or 1)
    I0, I1 = _cases.synth.make_pair(3840, 2160, seed=1234)
    kw = dict(nscales=6, warps=10, eps=0.001)
    u1, u2, iters, _ = gpu.Dual_TVL1_optic_flow_multiscale(I0, I1, **kw)
    r1, r2, riters, _ = cpu.multiscale(I0, I1, **kw)
    d = np.maximum(np.abs(u1 - r1), np.abs(u2 - r2))
    print("4K vs %s: mean|d|=%.3g max|d|=%.3g, %d px > 1e-2, iters per level=%s"
          % (kind, d.mean(), d.max(), int((d > MAX_TOL).sum()), iters.sum(axis=1).tolist()))
    assert np.array_equal(iters, riters), (iters.tolist(), riters.tolist())
    assert d.mean() <= MEAN_TOL
    assert d.max() <= MAX_TOL
    if available("reference", np.float32):
        c32 = CpuTvl1("reference", np.float32)
        c32.set_threads(os.cpu_count() or 1)
        f1, f2, fiters, _ = c32.multiscale(I0, I1, **kw)
        df = np.maximum(np.abs(f1 - r1), np.abs(f2 - r2))
        assert d.max() < df.max() and int((d > MAX_TOL).sum()) < int((df > MAX_TOL).sum())
