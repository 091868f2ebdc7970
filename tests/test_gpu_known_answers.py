"""The reference's own acceptance tests (feng/ddc/testing/test_ddc.py:22-332, test_cwg.py:15-92) restated against the
B200 module: same tones, same FFT length, same thresholds, same expected bins; the recorded answers of the reference
itself (tests/golden/meta.json:known_answers) are checked too."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from dc_sand_b200 import cwg, ddc  # noqa: E402

FS = 1712e6
FFT_LENGTH = 2**15


@pytest.fixture
def DDC_fixture(taps_dir):
    return ddc.DigitalDownConverter(decimation_factor=16, sampling_frequency=FS,
                                    ddc_coeff_filename=os.path.join(taps_dir, "ddc_coeff_107MHz.csv"))


def _tone(freq, n):
    return cwg.generate_carrier_wave(cw_scale=1, freq=freq, sampling_frequency=FS, num_samples=n, noise_scale=0,
                                     complex=False)


def _spectrum(DDC_fixture, freqs, mixing_freq):
    n = FFT_LENGTH * DDC_fixture.decimation_factor * 2
    data = sum(_tone(f, n) for f in freqs)
    decimated = DDC_fixture.run(data, mixing_freq)
    return np.abs(np.power(np.fft.fft(decimated[-FFT_LENGTH:], axis=-1), 2))


def _expected_bin(DDC_fixture, freq, mixing_freq):
    return int(np.floor((freq - mixing_freq) / ((FS / DDC_fixture.decimation_factor) / FFT_LENGTH)))


def test_run_ddc_center_cw(DDC_fixture, meta):
    p = _spectrum(DDC_fixture, [100e6], 100e6)
    bins = np.where(p > 1e5)[0]
    assert list(bins) == [_expected_bin(DDC_fixture, 100e6, 100e6)] == meta["known_answers"]["center"]["bins"]


def test_run_ddc_dual_cw(DDC_fixture, meta):
    p = _spectrum(DDC_fixture, [100e6, 103343750], 100e6)
    bins = list(np.where(p > 1e5)[0])
    assert bins == [0, _expected_bin(DDC_fixture, 103343750, 100e6)] == meta["known_answers"]["dual"]["bins"]


def test_run_ddc_bandedge_cw(DDC_fixture, meta):
    lo, hi = 51019287.109375, 148980712.890625
    p = _spectrum(DDC_fixture, [lo, hi], 100e6)
    bins = list(np.where(p > 1e5)[0])
    exp_hi = _expected_bin(DDC_fixture, hi, 100e6)
    exp_lo = FFT_LENGTH + _expected_bin(DDC_fixture, lo, 100e6)
    assert bins == sorted([exp_hi, exp_lo]) == meta["known_answers"]["bandedge"]["bins"] == [15000, 17768]


def test_run_ddc_out_of_band_cw(DDC_fixture, meta):
    p = _spectrum(DDC_fixture, [100e6, 214e6], 100e6)
    bins = list(np.where(p > 1e3)[0])
    assert bins == [0] == meta["known_answers"]["out_of_band"]["bins"]
    srt = np.sort(p)
    rejection_db = 10 * np.log10(srt[-1] / srt[-2])
    assert rejection_db > 60  # test_ddc.py:268,332
    assert abs(rejection_db - meta["known_answers"]["out_of_band"]["rejection_db"]) < 1.0
