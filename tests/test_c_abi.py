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


def test_reference_prototype_recipe(tmp_path):
    """oracle/build_ref.sh: a no-op where /root/reference is absent (the GPU box); here it compiles the reference's CUDA
    prototype for sm_100a into cubins that export `kernel_ddc` (timing baseline of tools/ref_gpu_prototype.py)."""
    script = os.path.join(ROOT, "oracle", "build_ref.sh")
    r = subprocess.run(["sh", script, str(tmp_path / "no_such_reference")], capture_output=True, text=True)
    assert r.returncode == 0 and "not present" in r.stdout
    if not (os.path.isdir("/root/reference") and shutil.which("nvcc") and shutil.which("cuobjdump")):
        pytest.skip("needs /root/reference and the CUDA toolkit")
    subprocess.run(["sh", script], check=True, capture_output=True)
    for name in ("kernel_ddc_2p14.cubin", "kernel_ddc_2p28.cubin"):
        path = os.path.join(ROOT, "oracle", "_ref", name)
        out = subprocess.run(["cuobjdump", "-elf", path], capture_output=True, text=True).stdout
        assert "_Z10kernel_ddcPfS_S_fS_S_" in out and "sm_100" in subprocess.run(
            ["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout + out, name
