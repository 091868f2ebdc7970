"""The boundary is a C ABI: a plain C99 program (tests/c_abi/ddc_c_client.c, compiled with gcc -std=c99 -pedantic against
include/ddcb200.h and linked to libddcb200.so) uses it without Python, C++ or torch."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "c_abi", "ddc_c_client.c")
LIBDIR = os.path.join(ROOT, "dc_sand_b200")


def _build(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = os.path.join(str(tmp_path), "ddc_c_client")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
                    "-o", exe, SRC, "-L", LIBDIR, "-l:libddcb200.so", "-lm", f"-Wl,-rpath,{LIBDIR}"], check=True)
    return exe


def test_c_client_compiles_links_and_runs_without_gpu(tmp_path):
    out = subprocess.run([_build(tmp_path), "--no-gpu"], check=True, capture_output=True, text=True).stdout
    assert "C ABI surface OK" in out


@pytest.mark.gpu
def test_c_client_runs_the_ddc(tmp_path):
    out = subprocess.run([_build(tmp_path)], check=True, capture_output=True, text=True).stdout
    assert "C ABI OK" in out and "fused" in out, out
