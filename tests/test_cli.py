"""The reference's UNMODIFIED command-line program (src/tvl1flow_main.cpp) linked against
libtvl1_b200.so (cli/tvl1flow) and, for A/B, against the reference's own CPU solver
(cli/tvl1flow_ref); image IO through cli/iio_lite.cpp.  The .flo file is the only on-disk contract
(src/iio.cpp:2754-2776)."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import _cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cli", "tvl1flow")
CLI_REF = os.path.join(ROOT, "cli", "tvl1flow_ref")


def write_pgm(path, img):
    img = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    with open(path, "wb") as f:
        f.write(b"P5\n# synthetic\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(img.tobytes())
    return img


def read_flo(path):
    with open(path, "rb") as f:
        assert f.read(4) == b"PIEH"
        w, h = struct.unpack("<ii", f.read(8))
        d = np.frombuffer(f.read(), np.float32).reshape(h, w, 2)
    return d[..., 0], d[..., 1]


def run_cli(exe, tmp, nx=96, ny=72, args=("0", "0.25", "0.15", "0.3", "3", "0.5", "3", "0.01", "1")):
    I0, I1 = _cases.synth.make_pair(nx, ny, seed=77, scale=0.4)
    q0 = write_pgm(tmp / "a.pgm", I0)
    q1 = write_pgm(tmp / "b.pgm", I1)
    out = tmp / ("out_%s.flo" % os.path.basename(exe))
    p = subprocess.run([exe, str(tmp / "a.pgm"), str(tmp / "b.pgm"), str(out), *args],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    iters = [int(m) for m in re.findall(r"Warping: \d+, Iterations: (\d+), Error:", p.stderr)]
    scales = re.findall(r"Scale (\d+): (\d+)x(\d+)", p.stderr)
    return q0, q1, read_flo(out), iters, scales, p.stderr


@pytest.mark.skipif(not os.path.exists(CLI_REF), reason="cli/tvl1flow_ref not built (needs /root/reference)")
def test_reference_cli_with_iio_lite_matches_oracle(tmp_path, oracle_f64):
    """CPU: PGM in -> reference solver -> .flo out, through our IO shim, equals the oracle on the
    same 8-bit images (validates iio_lite.cpp's readers and the .flo writer)."""
    q0, q1, (u1, u2), iters, scales, err = run_cli(CLI_REF, tmp_path)
    r1, r2, riters, _ = oracle_f64.multiscale(q0.astype(np.float64), q1.astype(np.float64),
                                              nscales=3, zfactor=0.5, warps=3, eps=0.01)
    assert iters == riters.ravel().tolist()
    assert [(int(a), int(b), int(c)) for a, b, c in scales] == [(2, 24, 18), (1, 48, 36), (0, 96, 72)]
    assert np.array_equal(u1, r1.astype(np.float32)) and np.array_equal(u2, r2.astype(np.float32))


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(CLI), reason="cli/tvl1flow not built")
def test_cli_drop_in_on_gpu(tmp_path, oracle_f64):
    """GPU: the same unmodified main() linked against libtvl1_b200.so; same verbose lines, same
    iteration counts, flow within the north_star tolerance."""
    q0, q1, (u1, u2), iters, scales, err = run_cli(CLI, tmp_path)
    r1, r2, riters, _ = oracle_f64.multiscale(q0.astype(np.float64), q1.astype(np.float64),
                                              nscales=3, zfactor=0.5, warps=3, eps=0.01)
    assert iters == riters.ravel().tolist(), err
    assert [(int(a), int(b), int(c)) for a, b, c in scales] == [(2, 24, 18), (1, 48, 36), (0, 96, 72)]
    d = np.concatenate([np.abs(u1 - r1).ravel(), np.abs(u2 - r2).ravel()])
    assert d.mean() <= 1e-3 and d.max() <= 1e-2
    # the CLI clamps nscales by image size (tvl1flow_main.cpp:185-188) and warns about bad values
    _, _, _, iters2, scales2, err2 = run_cli(CLI, tmp_path, args=("0", "0.9", "0.15", "0.3", "100", "0.5", "2", "0.01", "1"))
    assert "tau changed to 0.25" in err2
    assert len(scales2) == 3 and len(iters2) == 6


CLI_SEQ = os.path.join(ROOT, "cli", "tvl1flow_seq")


def test_sequence_cli_usage_errors():
    """CPU: argument handling of the video front end (no GPU work is reached)."""
    if not os.path.exists(CLI_SEQ):
        pytest.skip("cli/tvl1flow_seq not built")
    p = subprocess.run([CLI_SEQ], capture_output=True, text=True, timeout=60)
    assert p.returncode == 1 and "usage:" in p.stderr
    p = subprocess.run([CLI_SEQ, "-p", "a.pgm", "b.pgm", "c.pgm"], capture_output=True, text=True, timeout=60)
    assert p.returncode == 1 and "usage:" in p.stderr


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(CLI) and os.path.exists(CLI_SEQ)), reason="cli not built")
def test_sequence_cli_equals_pairwise_cli(tmp_path):
    """GPU: four frames through tvl1flow_seq give the same three .flo files, byte for byte (the
    reference main solves in double, the sequence front end in float: compare values, and iteration
    counts), as three runs of the reference's two-image program."""
    frames = []
    for k in range(4):
        I0, _ = _cases.synth.make_pair(96, 72, seed=5, scale=0.4)
        write_pgm(tmp_path / ("f%d.pgm" % k), np.roll(I0, (k, 2 * k), axis=(0, 1)))
        frames.append(str(tmp_path / ("f%d.pgm" % k)))
    p = subprocess.run([CLI_SEQ, "-o", str(tmp_path / "seq_"), "-s", "3", "-w", "3", "-v", "-b", "2", *frames],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    seq_iters = [int(m) for m in re.findall(r"Warping: \d+, Iterations: (\d+), Error:", p.stderr)]
    assert len(seq_iters) == 3 * 9
    for k in range(3):
        out = tmp_path / ("pair%d.flo" % k)
        q = subprocess.run([CLI, frames[k], frames[k + 1], str(out), "0", "0.25", "0.15", "0.3", "3", "0.5", "3",
                            "0.01", "1"], capture_output=True, text=True, timeout=300)
        assert q.returncode == 0, q.stderr
        iters = [int(m) for m in re.findall(r"Warping: \d+, Iterations: (\d+), Error:", q.stderr)]
        assert iters == seq_iters[9 * k:9 * k + 9]
        a1, a2 = read_flo(out)
        b1, b2 = read_flo(tmp_path / ("seq_%04d.flo" % k))
        assert np.array_equal(a1, b1) and np.array_equal(a2, b2)
    # independent pairs: (f0 f1)(f2 f3)
    p = subprocess.run([CLI_SEQ, "-p", "-o", str(tmp_path / "par_"), "-s", "3", "-w", "3", *frames],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    for k, src in ((0, 0), (1, 2)):
        a1, a2 = read_flo(tmp_path / ("par_%04d.flo" % k))
        b1, b2 = read_flo(tmp_path / ("seq_%04d.flo" % src))
        assert np.array_equal(a1, b1) and np.array_equal(a2, b2)


# ---- the reference's Horn-Schunck program (src/horn_schunck_pyramidal_main.cpp), unmodified -------
HS_CLI = os.path.join(ROOT, "cli", "horn_schunck_pyramidal")
HS_CLI_REF = os.path.join(ROOT, "cli", "horn_schunck_pyramidal_ref")
# argv of :88-105: outfile nproc alpha nscales zfactor nwarps TOL maxiter verbose
HS_ARGS = ("1", "7", "3", "0.5", "3", "0.001", "60", "1")


def run_hs_cli(exe, tmp, nx=96, ny=72, args=HS_ARGS):
    I1, I2 = _cases.synth.make_pair(nx, ny, seed=77, scale=0.4)
    q1 = write_pgm(tmp / "a.pgm", I1)
    q2 = write_pgm(tmp / "b.pgm", I2)
    out = tmp / ("hs_%s.flo" % os.path.basename(exe))
    p = subprocess.run([exe, str(tmp / "a.pgm"), str(tmp / "b.pgm"), str(out), *args],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    iters = [int(m) for m in re.findall(r"Iterations (\d+) \(", p.stderr)]
    scales = re.findall(r"Scale: (\d+) (\d+)x(\d+)", p.stderr)
    return q1, q2, read_flo(out), iters, scales, p.stderr


@pytest.mark.skipif(not os.path.exists(HS_CLI_REF), reason="cli/horn_schunck_pyramidal_ref not built")
def test_reference_hs_cli_with_iio_lite_matches_oracle(tmp_path, oracle_f64):
    """CPU: the reference's Horn-Schunck program (one thread: argv nproc = 1) through our IO shim equals
    the oracle on the same 8-bit images."""
    q1, q2, (u, v), iters, scales, err = run_hs_cli(HS_CLI_REF, tmp_path)
    ru, rv, rit, _ = oracle_f64.hs_multiscale(q1.astype(np.float64), q2.astype(np.float64), alpha=7.0, nscales=3,
                                              zfactor=0.5, warps=3, tol=1e-3, maxiter=60)
    assert iters == rit.ravel().tolist()
    assert [(int(a), int(b), int(c)) for a, b, c in scales] == [(2, 24, 18), (1, 48, 36), (0, 96, 72)]
    assert np.array_equal(u, ru.astype(np.float32)) and np.array_equal(v, rv.astype(np.float32))


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(HS_CLI), reason="cli/horn_schunck_pyramidal not built")
def test_hs_cli_drop_in_on_gpu(tmp_path, oracle_f64):
    """GPU: the same unmodified main() linked against libtvl1_b200.so: same verbose lines, same sweep
    counts, flow within the tolerance; and the CLI's nscales clamp (main.cpp:136-143) still applies."""
    q1, q2, (u, v), iters, scales, err = run_hs_cli(HS_CLI, tmp_path)
    ru, rv, rit, _ = oracle_f64.hs_multiscale(q1.astype(np.float64), q2.astype(np.float64), alpha=7.0, nscales=3,
                                              zfactor=0.5, warps=3, tol=1e-3, maxiter=60)
    assert iters == rit.ravel().tolist(), err
    assert [(int(a), int(b), int(c)) for a, b, c in scales] == [(2, 24, 18), (1, 48, 36), (0, 96, 72)]
    assert "Multiscale Horn-Schunck of a 96x72 pair" in err and "Single-scale Horn-Schunck of a 24x18 image" in err
    d = np.concatenate([np.abs(u - ru).ravel(), np.abs(v - rv).ravel()])
    assert d.mean() <= 1e-3 and d.max() <= 1e-2
    _, _, _, iters2, scales2, _ = run_hs_cli(HS_CLI, tmp_path, args=("1", "7", "10", "0.5", "2", "0.001", "20", "1"))
    assert len(scales2) == 3 and len(iters2) == 6       # 96x72: N = 1 + log2(120 / 16) = 3.9 -> 3 levels


# ---- TV-L1 with occlusions: the reference's unmodified src/tvl1occflow_main.cpp ---------------------------
OCC_CLI = os.path.join(ROOT, "cli", "tvl1occflow")
OCC_CLI_REF = os.path.join(ROOT, "cli", "tvl1occflow_ref")


def read_png_grey8(path):
    """Reads back the 8-bit grey PNG iio_lite.cpp writes (any zlib stream, filter 0 rows)."""
    import zlib
    b = open(path, "rb").read()
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(b):
        n, typ = struct.unpack(">I4s", b[pos:pos + 8])
        body = b[pos + 8:pos + 8 + n]
        assert zlib.crc32(typ + body) == struct.unpack(">I", b[pos + 8 + n:pos + 12 + n])[0]
        if typ == b"IHDR":
            w, h, depth, colour = struct.unpack(">IIBB", body[:10])
            assert (depth, colour) == (8, 0)
        elif typ == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, w + 1)
    assert not raw[:, 0].any()
    return raw[:, 1:]


def run_occ_cli(exe, tmp, nx=112, ny=80):
    I_1, I0, I1 = _cases.synth.make_triple(nx, ny, seed=31, scale=0.5)
    q = [write_pgm(tmp / ("f%d.pgm" % k), im) for k, im in enumerate((I_1, I0, I1))]
    flo, occ = tmp / ("occ_%s.flo" % os.path.basename(exe)), tmp / ("occ_%s.png" % os.path.basename(exe))
    # I_1 I0 I1 filtI0 out occ nproc lambda alpha beta theta nscales zfactor nwarps epsilon verbose
    args = [str(tmp / "f0.pgm"), str(tmp / "f1.pgm"), str(tmp / "f2.pgm"), str(tmp / "f1.pgm"), str(flo), str(occ),
            "1", "0.15", "0.01", "0.15", "0.3", "100", "0.5", "2", "0.01", "1"]
    p = subprocess.run([exe] + args, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    iters = [int(m) for m in re.findall(r"Warping: \d+, Iterations: (\d+), Error:", p.stderr)]
    return q, read_flo(flo), read_png_grey8(occ), iters, p.stderr


def occ_expected(q, nx=112, ny=80):
    from oracle.loader import CpuOcc
    P = CpuOcc("port", np.float64)
    P.set_threads(1)
    nscales = int(np.floor(np.log(np.float32(min(nx, ny)) / 16.0) / np.log(1. / 0.5))) + 1   # main.cpp:191-196
    return P.multiscale(q[0], q[1], q[2], None, nscales=nscales, zfactor=0.5, warps=2, eps=0.01), nscales


@pytest.mark.skipif(not os.path.exists(OCC_CLI_REF), reason="cli/tvl1occflow_ref not built (needs /root/reference)")
def test_reference_occ_cli_with_iio_lite_matches_oracle(tmp_path):
    """CPU: three PGM frames -> the reference's occlusion solver (zero-filling new[]) -> .flo + occlusion PNG
    through our IO shim: equal to the oracle on the same 8-bit frames."""
    q, (u1, u2), occ, iters, err = run_occ_cli(OCC_CLI_REF, tmp_path)
    (r1, r2, rchi, riters, _), nscales = occ_expected(q)
    assert nscales == 3 and iters == riters.ravel().tolist(), err
    assert np.array_equal(u1, r1.astype(np.float32)) and np.array_equal(u2, r2.astype(np.float32))
    assert np.array_equal(occ, (rchi * 255).astype(np.uint8))


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(OCC_CLI), reason="cli/tvl1occflow not built")
def test_occ_cli_drop_in_on_gpu(tmp_path):
    """GPU: the same unmodified main() linked against libtvl1_b200.so: same verbose lines and iteration
    counts, the .flo file and the occlusion map IDENTICAL to the reference's (fp64 path, bit-exact)."""
    q, (u1, u2), occ, iters, err = run_occ_cli(OCC_CLI, tmp_path)
    (r1, r2, rchi, riters, _), _ = occ_expected(q)
    assert iters == riters.ravel().tolist(), err
    assert np.array_equal(u1, r1.astype(np.float32)) and np.array_equal(u2, r2.astype(np.float32))
    assert np.array_equal(occ, (rchi * 255).astype(np.uint8))
    assert 0 < occ.astype(bool).sum() < occ.size
