"""Seeded inputs shared by the golden-vector generator, the oracle tests and the GPU parity
tests.  Everything here is deterministic numpy; nothing reads /root/reference."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_spec = importlib.util.spec_from_file_location(
    "_tvl1_synth", os.path.join(ROOT, "optical-flow-1_b200", "synth.py"))
synth = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(synth)

SIGMA_PRE = 0.8                                   # src/tvl1flow.cpp:23
SIGMA_ZOOM_HALF = 0.6 * np.sqrt(1.0 / 0.25 - 1.0)  # src/zoom.cpp:60 at factor 0.5

FUNC_SHAPE = (23, 31)   # (ny, nx): odd sizes, not multiples of 4


def func_inputs(shape=FUNC_SHAPE, seed=7):
    rs = np.random.RandomState(seed)
    ny, nx = shape
    return dict(
        I=rs.uniform(0, 255, shape),
        J=rs.uniform(10, 200, shape),
        u=rs.uniform(-5, 5, shape),
        v=rs.uniform(-5, 5, shape),
    )


def run_function_cases(cpu):
    """Outputs of every helper on the path for the seeded inputs; `cpu` is a CpuTvl1."""
    x = func_inputs()
    I, J, u, v = x["I"], x["J"], x["u"], x["v"]
    ny, nx = I.shape
    out = {}
    out["normalize0"], out["normalize1"] = cpu.normalize(I, J)
    out["gaussian_pre"] = cpu.gaussian(I, SIGMA_PRE)
    out["gaussian_zoom"] = cpu.gaussian(I, SIGMA_ZOOM_HALF)
    out["zoom_out_0.5"] = cpu.zoom_out(I, 0.5)
    out["zoom_out_0.7"] = cpu.zoom_out(I, 0.7)
    out["zoom_in"] = cpu.zoom_in(I, 2 * nx + 1, 2 * ny - 1)
    out["cgrad_x"], out["cgrad_y"] = cpu.centered_gradient(I)
    out["warp"] = cpu.warp(I, u, v, True)
    out["warp_clamped"] = cpu.warp(I, u, v, False)
    out["divergence"] = cpu.divergence(u, v)
    out["fgrad_x"], out["fgrad_y"] = cpu.forward_gradient(I)
    return out


SOLVER_CASES = {
    # name: (nx, ny, seed, motion scale, kwargs)
    "ms_64x48": dict(nx=64, ny=48, seed=1234, scale=0.5,
                     kw=dict(nscales=3, zfactor=0.5, warps=5, eps=0.01)),
    "ms_61x47_z07": dict(nx=61, ny=47, seed=99, scale=0.4,
                         kw=dict(nscales=3, zfactor=0.7, warps=3, eps=0.02)),
    "ms_96x64_cap": dict(nx=96, ny=64, seed=5, scale=0.5,
                         kw=dict(nscales=2, zfactor=0.5, warps=2, eps=0.0003)),
}


def solver_inputs(case):
    return synth.make_pair(case["nx"], case["ny"], seed=case["seed"], scale=case["scale"])


def run_solver_case(cpu, case):
    I0, I1 = solver_inputs(case)
    return cpu.multiscale(I0, I1, **case["kw"])


# ---- pyramidal Horn-Schunck (src/horn_schunck_pyramidal.cpp), SURVEY.md section 8f-4 ----------
HS_CASES = {
    # default alpha / TOL of src/horn_schunck_pyramidal_main.cpp:25-30; fewer warps / sweeps so
    # the fixtures stay small and fast
    "hs_64x48": dict(nx=64, ny=48, seed=1234, scale=0.5,
                     kw=dict(alpha=7.0, nscales=3, zfactor=0.5, warps=4, tol=1e-4, maxiter=60)),
    "hs_61x47_z07": dict(nx=61, ny=47, seed=99, scale=0.4,
                         kw=dict(alpha=15.0, nscales=3, zfactor=0.7, warps=3, tol=1e-3, maxiter=150)),
    # loose TOL: every warp step stops on the tolerance, not on maxiter
    "hs_96x64_tol": dict(nx=96, ny=64, seed=5, scale=0.5,
                         kw=dict(alpha=7.0, nscales=2, zfactor=0.5, warps=3, tol=2e-2, maxiter=150)),
}


def run_hs_case(cpu, case):
    I1, I2 = solver_inputs(case)
    return cpu.hs_multiscale(I1, I2, **case["kw"])


def hs_sor_inputs(nx=37, ny=29, seed=21):
    """A seeded SOR system in the shape src/horn_schunck_pyramidal.cpp:127-137 produces (warped
    gradients with exact zeros where the warp leaves the image, as border_out = true gives)."""
    rs = np.random.RandomState(seed)
    shape = (ny, nx)
    ix = rs.uniform(-20, 20, shape)
    iy = rs.uniform(-20, 20, shape)
    out = rs.uniform(0, 1, shape) < 0.08
    ix[out] = 0.0
    iy[out] = 0.0
    return dict(I1=rs.uniform(0, 255, shape), I2w=np.where(out, 0.0, rs.uniform(0, 255, shape)),
                I2wx=ix, I2wy=iy, u=rs.uniform(-3, 3, shape), v=rs.uniform(-3, 3, shape))


# ---- TV-L1 with occlusions: src/tvl1occflow.cpp -----------------------------------------
OCC_CASES = {
    # CLI defaults of src/tvl1occflow_constants.h:14-23 (nscales clamped by image size as the CLI does)
    "occ_64x48": dict(nx=64, ny=48, seed=1234, scale=0.5,
                      kw=dict(lam=0.15, alpha=0.01, beta=0.15, theta=0.3, nscales=2, zfactor=0.5, warps=2, eps=0.01)),
    "occ_96x80_tight": dict(nx=96, ny=80, seed=7, scale=0.5,
                            kw=dict(lam=0.15, alpha=0.01, beta=0.15, theta=0.3, nscales=3, zfactor=0.5, warps=3, eps=0.002)),
    "occ_61x47_z07": dict(nx=61, ny=47, seed=99, scale=0.4,
                          kw=dict(lam=0.2, alpha=0.02, beta=0.1, theta=0.25, nscales=2, zfactor=0.7, warps=2, eps=0.005)),
}


def occ_inputs(case):
    I_1, I0, I1 = synth.make_triple(case["nx"], case["ny"], seed=case["seed"], scale=case["scale"])
    return I_1, I0, I1


def run_occ_case(cpu, case):
    I_1, I0, I1 = occ_inputs(case)
    return cpu.multiscale(I_1, I0, I1, None, **case["kw"])
