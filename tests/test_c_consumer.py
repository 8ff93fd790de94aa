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


@pytest.mark.gpu
def test_c_program_solves_a_pair(tmp_path):
    p = subprocess.run([build(tmp_path)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
