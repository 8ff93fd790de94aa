"""Randomised differential tests on the GPU (fixed seeds, so the cases are the same on every run):

* TV-L1 against the fp64 oracle over random shapes (widths not divisible by 4 included), pyramid depths, zoom
  factors, warps, epsilons, lambda / theta: the north_star bar on every case -- mean |d| <= 1e-3 px,
  max |d| <= 1e-2 px, identical iteration counts (profiles/r2u_fuzz_tvl1.txt: 60 cases, worst max 4.5e-4);
* TV-L1 with occlusions against the oracle: identical bits on every case."""
import numpy as np
import pytest

import _cases
import optical_flow_1_b200 as pkg
from oracle.loader import CpuOcc, CpuTvl1

pytestmark = pytest.mark.gpu


def test_tvl1_random_configurations_meet_the_bar():
    rs = np.random.RandomState(2026)
    gpu = pkg.TVL1(device=0)
    cpu = CpuTvl1("port", np.float64)
    for c in range(30):
        nx, ny = int(rs.randint(48, 420)), int(rs.randint(40, 300))
        zf = float(rs.choice([0.5, 0.5, 0.6, 0.75]))
        nscales = int(rs.randint(1, 5))
        while nscales > 1 and min(nx, ny) * zf ** (nscales - 1) < 14:
            nscales -= 1
        kw = dict(tau=0.25, lam=float(rs.choice([0.1, 0.15, 0.3])), theta=float(rs.choice([0.2, 0.3, 0.5])), nscales=nscales,
                  zfactor=zf, warps=int(rs.randint(1, 5)), eps=float(rs.choice([0.05, 0.01, 0.01, 0.004])))
        I0, I1 = _cases.synth.make_pair(nx, ny, seed=int(rs.randint(1, 10 ** 6)), scale=float(rs.uniform(0.2, 1.0)))
        g = gpu.Dual_TVL1_optic_flow_multiscale(I0, I1, **kw)
        r = cpu.multiscale(I0.astype(np.float64), I1.astype(np.float64), **kw)
        d = np.concatenate([np.abs(g[0] - r[0]).ravel(), np.abs(g[1] - r[1]).ravel()])
        assert np.array_equal(g[2], r[2]), (c, nx, ny, kw, g[2].tolist(), r[2].tolist())
        assert d.mean() <= 1e-3 and d.max() <= 1e-2, (c, nx, ny, kw, d.mean(), d.max())
    gpu.close()


def test_occlusion_solver_random_configurations_are_bit_identical():
    rs = np.random.RandomState(77)
    gpu = pkg.TVL1Occ(device=0)
    cpu = CpuOcc("port", np.float64)
    for c in range(10):
        nx, ny = int(rs.randint(40, 150)), int(rs.randint(36, 110))
        zf = float(rs.choice([0.5, 0.5, 0.7]))
        nscales = int(rs.randint(1, 4))
        while nscales > 1 and nx * zf ** (nscales - 1) < 12:
            nscales -= 1
        kw = dict(lam=float(rs.choice([0.1, 0.15, 0.25])), alpha=float(rs.choice([0.005, 0.01, 0.03])),
                  beta=float(rs.choice([0.1, 0.15])), theta=float(rs.choice([0.25, 0.3])), nscales=nscales, zfactor=zf,
                  warps=int(rs.randint(1, 4)), eps=float(rs.choice([0.02, 0.01, 0.004])))
        I_1, I0, I1 = _cases.synth.make_triple(nx, ny, seed=int(rs.randint(1, 10 ** 6)), scale=float(rs.uniform(0.3, 0.8)))
        g = gpu.Dual_TVL1_optic_flow_multiscale(I_1, I0, I1, None, **kw)
        r = cpu.multiscale(I_1, I0, I1, None, **kw)
        assert np.array_equal(g[3], r[3]), (c, nx, ny, kw, g[3].tolist(), r[3].tolist())
        for k in range(3):
            assert np.array_equal(g[k], r[k]), (c, nx, ny, kw, k, float(np.abs(g[k] - r[k]).max()))
    gpu.close()
