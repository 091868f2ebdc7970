"""CPU test that pins the oracle to the LIVE reference: where /root/reference exists (the build container; it does not exist
on the GPU box, so the test skips itself there) the unmodified `DigitalDownConverter.run()` and `cwg.generate_carrier_wave`
of feng/ddc/src are imported and run on seeded random inputs, and the oracle's faithful restatement must reproduce them bit for
bit -- beyond the stored vectors of tests/golden/ (make_golden.py, make_golden_r2.py), which were produced the same way.

Reference lines exercised: feng/ddc/src/ddc.py:13-48 (constructor, genfromtxt), :121-162 (run), cwg.py:6-44."""
import importlib
import os
import sys

import numpy as np
import pytest

from oracle import ddc_oracle as orc
from dc_sand_b200 import taps as taps_mod

REF_SRC = "/root/reference/feng/ddc/src"
FS = 1712e6


@pytest.fixture(scope="module")
def ref():
    if not os.path.isfile(os.path.join(REF_SRC, "ddc.py")):
        pytest.skip("the reference tree is not on this machine")
    sys.dont_write_bytecode = True          # /root/reference is read-only
    sys.path.insert(0, REF_SRC)
    try:
        for name in ("ddc", "cwg"):
            sys.modules.pop(name, None)
        mods = importlib.import_module("ddc"), importlib.import_module("cwg")
    finally:
        sys.path.remove(REF_SRC)
    assert os.path.dirname(os.path.abspath(mods[0].__file__)) == REF_SRC
    yield mods
    for name in ("ddc", "cwg"):
        sys.modules.pop(name, None)


def test_oracle_reproduces_live_reference_on_random_inputs(ref):
    ddc_ref, _ = ref
    rng = np.random.default_rng(20261019)
    cases = [(16, "ddc_coeff_107MHz.csv"), (16, "ddc_coeff_53MHz.csv"), (32, "ddc_coeff_53MHz.csv"), (4, "ddc_coeff_107MHz.csv"),
             (7, "ddc_coeff_107MHz.csv"), (1, "ddc_coeff_107MHz.csv"), (64, "ddc_coeff_53MHz.csv"), (16, "ddc_coeff_107MHz.csv")]
    for i, (d, csv) in enumerate(cases):
        n = int(rng.integers(257, 30000)) if i else 100        # the first case is shorter than the filter (scipy swaps operands)
        fc = float(rng.uniform(1e6, 850e6)) * (-1.0 if i == 5 else 1.0)
        x = np.clip(np.rint(rng.normal(0.0, 150.0, n)), -512, 511).astype(np.float32)
        path = os.path.join(REF_SRC, csv)
        y_ref = ddc_ref.DigitalDownConverter(decimation_factor=d, sampling_frequency=int(FS), ddc_coeff_filename=path).run(x, fc)
        y = orc.ddc_reference(x, fc, taps_mod.coefficients(csv), d, FS)
        assert y.dtype == y_ref.dtype and y.shape == y_ref.shape == (orc.out_len(n, 256, d),), (i, y.shape, y_ref.shape)
        assert np.array_equal(y, y_ref), (i, d, csv, n, fc)
        if n >= 256:   # the O(window) float64 form the full-size GPU tests use
            w = orc.ddc_windowed_f64(x, 0, len(y), orc.phase_step_cycles(n, fc, FS), taps_mod.coefficients(csv), d)
            assert np.abs(w - y_ref).max() <= 1e-6 * np.abs(y_ref).max(), i


def test_shipped_taps_equal_the_reference_files(ref):
    """The coefficient tables regenerated from integer numerators (dc_sand_b200/taps.py) against genfromtxt of the reference's
    own CSV files, as the reference's constructor reads them (ddc.py:33-48)."""
    for csv in ("ddc_coeff_107MHz.csv", "ddc_coeff_53MHz.csv"):
        want = np.genfromtxt(os.path.join(REF_SRC, csv), delimiter=",")
        assert np.array_equal(taps_mod.coefficients(csv), want), csv


def test_carrier_wave_restatement_against_live_cwg(ref):
    _, cwg_ref = ref
    for freq, n, cplx in ((100e6, 4096, False), (214e6, 8192, True), (53.5e6, 1000, True)):
        want = cwg_ref.generate_carrier_wave(cw_scale=1, freq=freq, sampling_frequency=int(FS), num_samples=n, noise_scale=0.0,
                                             complex=cplx)
        got = orc.carrier_wave(1, freq, FS, n, complex=cplx)
        assert got.dtype == want.dtype and np.array_equal(got, want), (freq, n, cplx)
