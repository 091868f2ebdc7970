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
