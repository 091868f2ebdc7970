"""CPU tests: the oracle (oracle/ddc_oracle.py) is pinned against vectors produced by the unmodified reference
(tests/golden/make_golden.py), so it can stand in for the reference on the GPU box."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import ddc_oracle as orc
from dc_sand_b200 import taps as taps_mod


def _taps(csv):
    return taps_mod.coefficients(csv)


def _cases(meta):
    return [k for k in meta if k not in ("c1", "known_answers")]


def test_taps_regenerated_from_q17_match_oracle_helper(meta):
    import json, os
    from conftest import GOLDEN

    q = json.load(open(os.path.join(GOLDEN, "taps_q17.json")))
    for name in taps_mod.NAMES:
        t = _taps(name)
        assert len(t) == 256
        assert np.array_equal(t, orc.taps_from_q17(q[name]))
        assert np.array_equal(t, t[::-1])  # linear phase
        assert abs(t.sum() - {"ddc_coeff_107MHz.csv": 1.4127943532, "ddc_coeff_53MHz.csv": 1.412719596}[name]) < 1e-9


def test_faithful_restatement_is_bit_identical_to_reference(meta, golden_small):
    """ddc_reference() uses the same library calls as the reference, so it must reproduce run() bit for bit."""
    for name in _cases(meta):
        m = meta[name]
        x = golden_small[name + (":xf" if m.get("float_input") else ":x")]
        y_ref = golden_small[name + ":y"]
        y = orc.ddc_reference(x.astype(np.float32), m["fc"], _taps(m["csv"]), m["d"], m["fs"])
        assert y.dtype == np.complex128
        assert len(y) == m["m"] == orc.out_len(m["n"], 256, m["d"])
        assert np.array_equal(y, y_ref), name


def test_faithful_noise_draw_does_not_change_result(meta, golden_small):
    m = meta["n16384_d16"]
    x = golden_small["n16384_d16:x"].astype(np.float32)
    y = orc.ddc_reference(x, m["fc"], _taps(m["csv"]), m["d"], m["fs"], faithful_noise=True)
    assert np.array_equal(y, golden_small["n16384_d16:y"])


def test_windowed_f64_matches_reference(meta, golden_small):
    """The O(window) float64 form used for 2^28-sample checks agrees with run() to 1e-12."""
    for name in _cases(meta):
        m = meta[name]
        if m["n"] < 256:
            continue  # swapped-operand corner is not a windowed case
        x = golden_small[name + (":xf" if m.get("float_input") else ":x")].astype(np.float32)
        step = orc.phase_step_cycles(m["n"], m["fc"], m["fs"])
        y = orc.ddc_windowed_f64(x, 0, m["m"], step, _taps(m["csv"]), m["d"])
        emax, el2 = rel_err(y, golden_small[name + ":y"])
        assert emax < 1e-12 and el2 < 1e-12, (name, emax, el2)
        # a window in the middle, fed from a slice of the input
        if m["m"] > 40:
            d, t = m["d"], 256
            lo = 17 * d
            hi = (17 + 20 - 1) * d + t
            yw = orc.ddc_windowed_f64(x[lo:hi], 17, 20, step, _taps(m["csv"]), d, x_base=lo)
            emax, _ = rel_err(yw, golden_small[name + ":y"][17:37])
            assert emax < 1e-12


def test_windowed_chunk_offsets_reproduce_one_shot(meta, golden_small):
    """sample_offset + whole-stream step reproduce the one-shot call when a stream is cut into chunks."""
    m = meta["n40000_d16_fc214"]
    x = golden_small["n40000_d16_fc214:x"].astype(np.float32)
    step = orc.phase_step_cycles(m["n"], m["fc"], m["fs"])
    d = m["d"]
    m0 = 1000
    chunk = x[m0 * d :]
    y = orc.ddc_windowed_f64(chunk, 0, m["m"] - m0, step, _taps(m["csv"]), d, sample_offset=m0 * d)
    emax, _ = rel_err(y, golden_small["n40000_d16_fc214:y"][m0:])
    assert emax < 1e-12


def test_c1_subsample(meta, golden_c1):
    """BASELINE config 1 (N = 2^20): oracle windows against the stored strided subsample of the reference output."""
    from dc_sand_b200 import synth

    m = meta["c1"]
    x = synth.digitiser_stream(m["n"], m["seed"]).astype(np.float32)
    step = orc.phase_step_cycles(m["n"], m["fc"], m["fs"])
    t = _taps(m["csv"])
    head = orc.ddc_windowed_f64(x, 0, 256, step, t, m["d"])
    tail = orc.ddc_windowed_f64(x, m["m"] - 256, 256, step, t, m["d"])
    assert rel_err(head, golden_c1["y_head"])[0] < 1e-12
    assert rel_err(tail, golden_c1["y_tail"])[0] < 1e-12
    idx = np.arange(0, m["m"], m["stride"])
    sub = np.array([orc.ddc_windowed_f64(x, int(i), 1, step, t, m["d"])[0] for i in idx[:200]])
    assert rel_err(sub, golden_c1["y_sub"][:200])[0] < 1e-12


def test_nco_phase_law_quirk():
    """The NCO step is int(N fc / fs) / (N - 1), not fc / fs (cwg.py:31-33)."""
    n, fc, fs = 1 << 20, 100e6, 1712e6
    step = orc.phase_step_cycles(n, fc, fs)
    assert step == int(n / (fs / fc)) / (n - 1)
    assert abs(step - fc / fs) > 1e-8
    with pytest.raises(ZeroDivisionError):
        orc.phase_step_cycles(n, 0.0, fs)


def test_cwg_restatement(golden_cwg, meta):
    cw = orc.carrier_wave(1, 100e6, 1712e6, 8192, complex=False)
    assert cw.dtype == np.float32 and np.array_equal(cw, golden_cwg["cwg_real_8192"])
    cwc = orc.carrier_wave(1, 214e6, 1712e6, 8192, complex=True)
    assert cwc.dtype == np.complex64 and np.array_equal(cwc, golden_cwg["cwg_complex_8192"])
    f = np.fft.rfft(cw)
    assert int(np.where(f == np.max(f))[0][0]) == meta["known_answers"]["cwg_real_bin"] == 478
    f = np.fft.fft(cwc)
    assert int(np.where(f == np.max(f))[0][0]) == meta["known_answers"]["cwg_complex_bin"] == 7168


def test_empty_input_raises():
    with pytest.raises(ValueError, match="Too few samples"):
        orc.ddc_reference(np.zeros(0, np.float32), 100e6, _taps("ddc_coeff_107MHz.csv"), 16, 1712e6)


def test_pack10_roundtrip_all_codes_all_phases():
    codes = np.arange(-512, 512, dtype=np.int16)
    for phase in range(4):
        s = np.zeros(4 * 1024, dtype=np.int16)
        s[phase::4] = codes
        s[(phase + 1) % 4 :: 4] = codes[::-1]
        p = orc.pack10(s)
        assert p.dtype == np.uint8 and len(p) == len(s) // 4 * 5
        assert np.array_equal(orc.unpack10(p), s)
    # bit layout: first sample occupies the top 10 bits of the first 2 bytes
    assert list(orc.pack10(np.array([-512, 0, 0, 0]))) == [0x80, 0, 0, 0, 0]
    assert list(orc.pack10(np.array([0, 0, 0, 1]))) == [0, 0, 0, 0, 1]
    assert list(orc.pack10(np.array([1, 0, 0, 0]))) == [0, 0x40, 0, 0, 0]
    rng = np.random.default_rng(0)
    s = rng.integers(-512, 512, size=4096).astype(np.int16)
    assert np.array_equal(orc.unpack10(orc.pack10(s)), s)


def test_oracle_is_bit_identical_on_round2_reference_vectors(meta_r2, golden_r2):
    """Input dtypes and argument corners beyond float32 (int16, int8, float64, complex input, negative centre frequency,
    full-range uniform codes): vectors produced by the unmodified reference (tests/golden/make_golden_r2.py)."""
    for name, m in meta_r2.items():
        x = golden_r2[name + ":x"]
        assert str(x.dtype) == m["in_dtype"]
        y = orc.ddc_reference(x, m["fc"], _taps(m["csv"]), m["d"], m["fs"])
        assert y.dtype == np.complex128 and np.array_equal(y, golden_r2[name + ":y"]), name
