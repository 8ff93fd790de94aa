"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built by
`make -C oracle ref` from /root/reference/src).  Run in the build container only:

    python tests/golden/make_golden.py

Inputs are regenerated from seeds by tests/_cases.py, so the fixtures hold outputs only.
Single-threaded so that the OpenMP error reduction (src/tvl1flow.cpp:151) is order-stable.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.loader import CpuOcc, CpuTvl1, build  # noqa: E402
import _cases  # noqa: E402


def main():
    build(ref=True)
    out = {}
    for dt in (np.float64, np.float32):
        R = CpuTvl1("reference", dt)
        R.set_threads(1)
        tag = "f64" if dt == np.float64 else "f32"
        out.update({"%s/%s" % (tag, k): v for k, v in _cases.run_function_cases(R).items()})
        for name, case in _cases.SOLVER_CASES.items():
            u1, u2, iters, errs = _cases.run_solver_case(R, case)
            out["%s/%s/u1" % (tag, name)] = u1
            out["%s/%s/u2" % (tag, name)] = u2
            out["%s/%s/iters" % (tag, name)] = iters
            out["%s/%s/errs" % (tag, name)] = errs
    path = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")

    # pyramidal Horn-Schunck (src/horn_schunck_pyramidal.cpp), one thread: the reference's SOR sweep
    # is an in-place update under an OpenMP parallel-for and only well defined then
    hs = {}
    for dt in (np.float64, np.float32):
        R = CpuTvl1("reference", dt)
        tag = "f64" if dt == np.float64 else "f32"
        for name, case in _cases.HS_CASES.items():
            u, v, iters, errs = _cases.run_hs_case(R, case)
            hs["%s/%s/u" % (tag, name)] = u
            hs["%s/%s/v" % (tag, name)] = v
            hs["%s/%s/iters" % (tag, name)] = iters
            hs["%s/%s/errs" % (tag, name)] = errs
    path = os.path.join(ROOT, "tests", "golden", "hs_reference_vectors.npz")
    np.savez_compressed(path, **hs)
    print("wrote", path, os.path.getsize(path), "bytes,", len(hs), "arrays")

    # TV-L1 with occlusions (src/tvl1occflow.cpp), reference built with the zero-filling new[] of
    # oracle/occ_ref_shim.cpp (the one defined reading of its uninitialised eta1/eta2)
    occ = {}
    for dt in (np.float64, np.float32):
        R = CpuOcc("reference", dt)
        R.set_threads(1)
        tag = "f64" if dt == np.float64 else "f32"
        for name, case in _cases.OCC_CASES.items():
            u1, u2, chi, iters, errs = _cases.run_occ_case(R, case)
            occ["%s/%s/u1" % (tag, name)] = u1
            occ["%s/%s/u2" % (tag, name)] = u2
            occ["%s/%s/chi" % (tag, name)] = chi.astype(np.uint8)
            occ["%s/%s/iters" % (tag, name)] = iters
    path = os.path.join(ROOT, "tests", "golden", "occ_reference_vectors.npz")
    np.savez_compressed(path, **occ)
    print("wrote", path, os.path.getsize(path), "bytes,", len(occ), "arrays")


if __name__ == "__main__":
    main()
