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
