"""A plain C99 program against include/tvl1_b200.h + libtvl1_b200.so (no Python in the data path)."""
import os
import subprocess

import pytest

import optical_flow_1_b200 as pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "abi_smoke.c")


def build(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(pkg.library_path())
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    SRC, "-o", exe, "-L", libdir, "-ltvl1_b200", "-Wl,-rpath," + libdir, "-lm"], check=True)
    return exe


def test_header_is_c99_and_library_links(tmp_path):
    exe = build(tmp_path)
    import torch
    if not torch.cuda.is_available():
        p = subprocess.run([exe], capture_output=True, text=True)
        assert p.returncode == 3 and "no CPU fallback" in p.stdout      # fails loudly without a GPU


def test_all_three_headers_are_c99(tmp_path):
    """tvl1_b200.h, hs_b200.h and occ_b200.h included from one strict C99 translation unit that links and calls
    the parameter helpers (no device needed)."""
    src = tmp_path / "hdr.c"
    src.write_text('#include "tvl1_b200.h"\n#include "hs_b200.h"\n#include "occ_b200.h"\n'
                   'int main(void) { occ_params p; hs_params h; tvl1_params t; occ_default_params(&p); hs_default_params(&h);\n'
                   '  tvl1_default_params(&t);\n'
                   '  return (occ_clamp_nscales(640, 480, p.nscales, p.zfactor) == 5 && h.warps == 10 && t.warps == 5) ? 0 : 1; }\n')
    exe = str(tmp_path / "hdr")
    libdir = os.path.dirname(pkg.library_path())
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", exe, "-L", libdir, "-ltvl1_b200", "-Wl,-rpath," + libdir, "-lm"], check=True)
    assert subprocess.run([exe]).returncode == 0


@pytest.mark.gpu
def test_c_program_solves_a_pair(tmp_path):
    p = subprocess.run([build(tmp_path)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
