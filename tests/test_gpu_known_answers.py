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


def test_cwg_on_device_matches_reference_vectors(golden_cwg):
    """ddcb200_cwg against the reference's own cwg output: float32 sin/cos of an exact fixed-point phase."""
    from dc_sand_b200 import cwg as mycwg

    cw = mycwg.generate_carrier_wave_gpu(cw_scale=1, freq=100e6, sampling_frequency=1712e6, num_samples=8192, noise_scale=0,
                                         complex=False)
    assert cw.dtype == torch.float32 and cw.shape == (8192,)
    assert np.abs(cw.cpu().numpy() - golden_cwg["cwg_real_8192"]).max() <= 3e-7
    cwc = mycwg.generate_carrier_wave_gpu(cw_scale=1, freq=214e6, sampling_frequency=1712e6, num_samples=8192, noise_scale=0,
                                          complex=True)
    assert cwc.dtype == torch.complex64
    assert np.abs(cwc.cpu().numpy() - golden_cwg["cwg_complex_8192"]).max() <= 3e-7
    # a later piece of the same wave continues the phase (sample_offset), as DDCStream needs
    tail = mycwg.generate_carrier_wave_gpu(1, 214e6, 1712e6, 4096, 0, True, sample_offset=4096, total_samples=8192)
    assert np.abs(tail.cpu().numpy() - golden_cwg["cwg_complex_8192"][4096:]).max() <= 3e-7


def test_cwg_on_device_noise_statistics_and_reproducibility():
    from dc_sand_b200 import cwg as mycwg

    n = 1 << 20
    a = mycwg.generate_carrier_wave_gpu(0.0, 100e6, 1712e6, n, 1.0, False, seed=7, n_streams=2)
    b = mycwg.generate_carrier_wave_gpu(0.0, 100e6, 1712e6, n, 1.0, False, seed=7, n_streams=2)
    c = mycwg.generate_carrier_wave_gpu(0.0, 100e6, 1712e6, n, 1.0, False, seed=8, n_streams=2)
    assert torch.equal(a, b) and not torch.equal(a, c) and not torch.equal(a[0], a[1])
    x = a[0].double().cpu().numpy()
    # truncated normal on [-1, 1] with sigma 0.5: std 0.4398, mean 0 (cwg.py:62-70)
    assert x.min() >= -1 and x.max() <= 1 and abs(x.mean()) < 3e-3 and abs(x.std() - 0.4398) < 3e-3
    # digitiser model: integers in the 10-bit range
    d = mycwg.generate_carrier_wave_gpu(100.0, 103.3e6, 1712e6, n, 40.0, False, seed=1, digitise=True).cpu().numpy()
    assert np.array_equal(d, np.rint(d)) and d.min() >= -512 and d.max() <= 511 and 70 < d.std() < 90


def test_reference_acceptance_tests_without_leaving_the_device(DDC_fixture, meta):
    """the reference's spectral acceptance tests (test_ddc.py:61-332) with tones generated in HBM
    (ddcb200_cwg), the fused DDC on device tensors and the spectrum check on the device -- only bin indices reach the host."""
    from dc_sand_b200 import cwg as mycwg, selfcheck

    n = FFT_LENGTH * DDC_fixture.decimation_factor * 2
    ka = meta["known_answers"]

    def bins(freqs, mix):
        data = sum(mycwg.generate_carrier_wave_gpu(1, f, FS, n, 0, False) for f in freqs)
        return selfcheck.spectrum_bins_above(DDC_fixture.run_tensor(data, mix), FFT_LENGTH, 1e5)

    assert bins([100e6], 100e6) == ka["center"]["bins"]
    assert bins([100e6, 103343750], 100e6) == ka["dual"]["bins"]
    assert bins([51019287.109375, 148980712.890625], 100e6) == ka["bandedge"]["bins"] == [15000, 17768]
    # out-of-band tone next to an in-band one: only the in-band bin above 1e3, second strongest bin > 60 dB down
    # (test_ddc.py:245-332)
    data = sum(mycwg.generate_carrier_wave_gpu(1, f, FS, n, 0, False) for f in (100e6, 214e6))
    y = DDC_fixture.run_tensor(data, 100e6)
    assert selfcheck.spectrum_bins_above(y, FFT_LENGTH, 1e3) == ka["out_of_band"]["bins"] == [0]
    top2 = torch.topk(selfcheck.power_spectrum(y, FFT_LENGTH), 2).values
    assert float(10 * torch.log10(top2[0] / top2[1])) > 60
    # and the device-resident path agrees with the host path sample for sample
    tone = mycwg.generate_carrier_wave_gpu(1, 103e6, FS, 1 << 20, 0, False)
    y = DDC_fixture.run_tensor(tone, 100e6).cpu().numpy()
    ref = DDC_fixture.run(tone.cpu().numpy(), 100e6)
    assert np.abs(y - ref).max() <= 1e-5 * np.abs(ref).max()
