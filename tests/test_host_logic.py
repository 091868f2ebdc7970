"""CPU tests of the host-side mirror of the reference interface and of the C-ABI surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from dc_sand_b200 import _lib, cwg, synth, taps
from oracle import ddc_oracle as orc


def test_header_symbols_are_exported_and_bound():
    """Every function include/ddcb200.h declares is exported by libddcb200.so and has a ctypes signature."""
    hdr = open(os.path.join(ROOT, "include", "ddcb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ddcb200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ddcb200_version() == 100


def test_out_len_matches_reference_lengths(meta):
    lib = _lib.load()
    for name, m in meta.items():
        if name in ("known_answers",):
            continue
        assert lib.ddcb200_out_len(m["n"], 256, m["d"]) == m["m"], name
    assert lib.ddcb200_out_len(0, 256, 16) == 0
    assert lib.ddcb200_out_len(1 << 28, 256, 16) == 16777201


def _plan(t, d, packed=0, aligned=1, variant=0, packed_engine=1):
    buf = ctypes.create_string_buffer(64)
    jt = _lib.load().ddcb200_plan(t, d, packed, aligned, variant, packed_engine, buf, 64)
    return buf.value.decode(), jt


def test_kernel_selection_table():
    """The dispatcher's selection is a pure function (ddcb200_plan): the BASELINE configs[3] grid (taps 64-1024 x decimation
    4-64) maps onto the kernel families DESIGN.md section 4 names, for float32 and for packed input on both engines, and the
    corners fall through to the tile / generic kernels."""
    taps_axis = (64, 128, 256, 512, 1024)
    f32 = {
        4: [("ws", 16), ("ws", 32), ("ws", 64), ("ws", 128), ("ws", 256)],
        8: [("ws", 8), ("ws", 16), ("ws", 32), ("ws", 64), ("ws", 128)],
        16: [("ws", 8), ("ws", 8), ("w", 16), ("w", 32), ("w", 64)],           # HBM-bound filters on whole-row tiles, then the fast FIR
        32: [("pd", 4), ("pd", 4), ("pd", 8), ("ws", 16), ("ws", 32)],         # phase-major while HBM-bound, sliced when FP32-bound
        64: [("pd", 4), ("pd", 4), ("pd", 4), ("pd", 8), ("ws", 16)],
    }
    for d, row in f32.items():
        assert [_plan(t, d) for t in taps_axis] == row, d
        # packed input: the tensor engine in every cell; on the CUDA cores the unpack stage + the float32 family of the cell,
        # except the one cell with a fused-unpack CUDA-core kernel (D = 16, 129 .. 256 taps)
        assert [_plan(t, d, packed=1)[0] for t in taps_axis] == ["tensor10"] * 5, d
        want = [("unpack+f32:" + f, jt) for f, jt in row]
        if d == 16:
            want[2] = ("w10s", 16)
        assert [_plan(t, d, packed=1, packed_engine=0) for t in taps_axis] == want, d
    assert _plan(256, 16) == ("w", 16)                            # the headline configuration
    assert _plan(200, 16) == ("w", 16) and _plan(2, 16) == ("w", 4)
    assert _plan(300, 8) == ("tile", 40)                          # padding to 48 blocks would cost more than 12 %
    assert _plan(2048, 4) == ("tile", 512)
    assert _plan(256, 3) == ("generic", 0)                        # odd decimation
    assert _plan(4096, 16) == ("generic", 0)                      # more taps than any fused kernel holds
    assert _plan(256, 16, aligned=0) == ("generic", 0)            # rows not 16-byte aligned
    assert _plan(256, 16, variant=1) == ("generic", 0)
    assert _plan(256, 16, packed=1, aligned=0) == ("unpack+f32:w", 16)
    assert _plan(256, 12, packed=1) == ("unpack+f32:generic", 0)
    assert _plan(256, 16, packed=1, variant=13, packed_engine=0) == ("tensor10", 0)   # forced
    assert _plan(256, 32, variant=8) == ("pd", 8) and _plan(256, 16, variant=8) == ("w", 16)   # a variant that is not built falls through
    buf = ctypes.create_string_buffer(8)
    assert _lib.load().ddcb200_plan(256, 16, 0, 1, 0, 1, buf, 8) == _lib.EINVAL


def test_tensor_engine_pipeline_sizing_rules():
    """The tensor-core engine's shared-memory pipeline (ddcb200_tensor_engine_geometry, csrc/k_tc.cu) over every filter length
    1 .. 2100 at every decimation it is built for.  The kernel's hand-shakes rest on these rules (csrc/ddc_kernel_tc.cuh):
    a team of unpack warps per sample stage AND per raw slot at most (a parity wait two phases ahead of its mbarrier would
    alias), the unpack warps cover every 16-sample group of a tile, K is a whole number of MMAs and holds the row plus the
    filter's overhang, the accumulator pair fits tensor memory, everything fits 227 KB and the 14-bit descriptor fields."""
    lib = _lib.load()
    out = (ctypes.c_int32 * 12)()
    assert lib.ddcb200_tensor_engine_geometry(256, 16, None) == _lib.EINVAL
    assert lib.ddcb200_tensor_engine_geometry(256, 12, out) == 0 and lib.ddcb200_tensor_engine_geometry(0, 16, out) == 0
    fits = {}
    for d in (4, 8, 16, 32, 64):
        last = 0
        for t in range(1, 2101):
            if not lib.ddcb200_tensor_engine_geometry(t, d, out):
                continue
            row_s, n, k, n_a, n_raw, teams, smem, groups, cap, pitch, raw_bytes, b_bytes = list(out)
            last = t
            assert row_s in (64, 128) and row_s % d == 0
            assert n == max(16, 4 * row_s // d) and n % 16 == 0 and 2 * n <= 512          # two accumulators in tensor memory
            assert k % 16 == 0 and row_s - d + t <= k < row_s - d + t + 16
            assert 2 <= n_a <= 3 and 2 <= n_raw <= 8
            assert teams in (2, 3) and teams <= n_a and teams <= n_raw and 12 % teams == 0
            assert groups * 16 == 128 * row_s + k - row_s and groups <= cap
            assert raw_bytes % 16 == 0 and raw_bytes >= 20 * groups                         # bulk copies: whole 16-byte pieces
            assert pitch % 16 == 0 and (pitch // 16) % 2 == 1                                # odd unit pitch: conflict-free stores
            assert pitch * (row_s // 8) >= 2 * groups * 16                                   # a stage holds the tile's fp16 samples
            assert b_bytes == n * k * 2
            assert smem <= 227 * 1024 and smem // 16 < (1 << 14)
            assert smem >= 1024 + b_bytes + n_a * pitch * (row_s // 8) + n_raw * raw_bytes
        fits[d] = last
    # the BASELINE sweep (taps 64 .. 1024 x decimation 4 .. 64) lies inside the engine at every decimation; configs[2] gets
    # 128-sample rows, three sample stages and three teams
    assert all(v >= 1024 for v in fits.values()), fits
    assert lib.ddcb200_tensor_engine_geometry(256, 16, out) == 1
    assert list(out)[:6] == [128, 32, 368, 3, 5, 3]
    # where the geometry says no, the dispatcher goes to the CUDA cores
    for t, d in ((fits[4] + 1, 4), (2100, 4)):
        assert lib.ddcb200_tensor_engine_geometry(t, d, out) == 0 and _plan(t, d, packed=1)[0].startswith("unpack+f32")


def test_host_thread_pool(tmp_path):
    """The pool of parked host threads that stages pageable input and widens complex128 output (csrc/host_pool.h): a C++
    stress test, under ThreadSanitizer when g++ links it."""
    import shutil
    import subprocess

    if shutil.which("g++") is None:
        pytest.skip("no g++")
    src = os.path.join(ROOT, "tests", "native", "host_pool_test.cpp")
    inc = os.path.join(ROOT, "dc_sand_b200", "csrc")
    exe = os.path.join(str(tmp_path), "host_pool_test")
    base = ["g++", "-std=c++17", "-O1", "-g", "-pthread", "-Wall", "-Werror", "-I", inc, src, "-o", exe]
    if subprocess.run(base + ["-fsanitize=thread"], capture_output=True).returncode != 0:
        subprocess.run(base, check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "host pool OK" in out.stdout, out.stdout + out.stderr


def test_library_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = ctypes.c_void_p()
    t = np.ones(8)
    rc = lib.ddcb200_create(ctypes.byref(h), 0, t.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 8, 2)
    assert rc == _lib.ECUDA and _lib.last_error()
    from dc_sand_b200 import DigitalDownConverter

    path = taps.write_csv("ddc_coeff_107MHz.csv", os.environ.get("TMPDIR", "/tmp"))
    ddc = DigitalDownConverter(16, 1712e6, path)
    with pytest.raises(RuntimeError):
        ddc.run(np.zeros(4096, np.float32), 100e6)  # no CPU fallback


def test_constructor_mirrors_reference_attributes(taps_dir):
    from dc_sand_b200 import DigitalDownConverter

    d = DigitalDownConverter(decimation_factor=16, sampling_frequency=1712e6,
                             ddc_coeff_filename=os.path.join(taps_dir, "ddc_coeff_107MHz.csv"))
    assert d.decimation_factor == 16 and d.sampling_frequency == 1712e6
    assert d.ddc_filter_coeffs.dtype == np.float64 and len(d.ddc_filter_coeffs) == 256
    assert np.array_equal(d.ddc_filter_coeffs, taps.coefficients("ddc_coeff_107MHz.csv"))
    with pytest.raises(ValueError, match="Too few samples in input data. Received 0"):
        d.run(np.zeros(0, np.float32), 100e6)
    with pytest.raises(ValueError):
        d.run(np.zeros((2, 4096), np.float32), 100e6)


def test_cwg_mirror_matches_reference_vectors(golden_cwg):
    cw = cwg.generate_carrier_wave(cw_scale=1, freq=100e6, sampling_frequency=1712e6, num_samples=8192, noise_scale=0,
                                   complex=False)
    assert cw.dtype == np.float32 and np.array_equal(cw, golden_cwg["cwg_real_8192"])
    cwc = cwg.generate_carrier_wave(cw_scale=1, freq=214e6, sampling_frequency=1712e6, num_samples=8192, noise_scale=0,
                                    complex=True)
    assert cwc.dtype == np.complex64 and np.array_equal(cwc, golden_cwg["cwg_complex_8192"])
    assert cwg.phase_step_cycles(1 << 20, 100e6, 1712e6) == orc.phase_step_cycles(1 << 20, 100e6, 1712e6)
    noisy = cwg.generate_carrier_wave(1, 100e6, 1712e6, 4096, 0.1, False)
    assert noisy.dtype == np.float32 and np.abs(noisy - cw[:4096] * 0 - noisy).max() == 0
    n = cwg._generate_noise(1.0, 20000, np.random.default_rng(1))
    assert n.dtype == np.float32 and n.min() >= -1 and n.max() <= 1 and 0.4 < n.std() < 0.5


def test_synth_packer_matches_oracle_packer():
    s = synth.digitiser_stream(4096, 99)
    assert s.dtype == np.int16 and s.min() >= -512 and s.max() <= 511
    assert np.array_equal(synth.pack10(s), orc.pack10(s))
    f = synth.digitiser_stream_fast(3 * (1 << 12) + 5, 7, block=1 << 12)
    assert len(f) == 3 * (1 << 12) + 5 and f.min() >= -512


def test_csv_text_round_trip(taps_dir):
    for name in taps.NAMES:
        parsed = np.genfromtxt(os.path.join(taps_dir, name), delimiter=",")
        assert np.array_equal(parsed, taps.coefficients(name))
