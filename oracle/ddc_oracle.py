"""CPU oracle for the feng/ddc digital down-converter.  TEST INFRASTRUCTURE ONLY.

This file restates, in NumPy/SciPy, the algorithm of the reference NumPy DDC so that the CUDA path can be
checked on machines where /root/reference does not exist (the GPU box).  Nothing in the product package
`dc_sand_b200/` imports it; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs do, and only as the checker or as the timed CPU baseline.

Parity status: PINNED.  `tests/golden/make_golden.py` imported the unmodified reference
(`/root/reference/feng/ddc/src/{ddc,cwg}.py`) in the build container and stored its `run()` outputs in
`tests/golden/*.npz`; `tests/test_oracle.py` checks every function here against those vectors (the faithful
restatement is bit-identical, the windowed float64 form agrees to <= 1e-12 relative).

Reference lines followed (paths relative to /root/reference):
  * NCO phase law                 feng/ddc/src/cwg.py:31-36   (samples_per_cycle, int() truncation, linspace, complex64)
  * zero-scaled noise term        feng/ddc/src/cwg.py:39-42,47-70 (adds exactly 0 inside run(); see `faithful_noise`)
  * mixer                         feng/ddc/src/ddc.py:51-66   (float32 * complex64 -> complex64)
  * FIR + normalisation           feng/ddc/src/ddc.py:85-100  (scipy.signal.convolve mode="valid", / sum(taps))
  * decimation                    feng/ddc/src/ddc.py:102-119 (filtered[0::D])
  * orchestration + empty check   feng/ddc/src/ddc.py:121-162
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "phase_step_cycles",
    "nco",
    "out_len",
    "ddc_reference",
    "ddc_windowed_f64",
    "ddc_ideal_f64",
    "pack10",
    "unpack10",
    "carrier_wave",
    "taps_from_q17",
]


# ----------------------------------------------------------------------------------------------------------
# NCO
# ----------------------------------------------------------------------------------------------------------
def phase_step_cycles(num_samples: int, center_freq: float, sampling_frequency: float) -> float:
    """Cycles per sample of the reference NCO (cwg.py:31-33).

    The reference builds the phase ramp with ``np.linspace(0, cycles, N)`` where
    ``cycles = int(N / (fs / fc))``; element n is therefore ``n * cycles / (N - 1)`` -- NOT ``n * fc / fs``.
    Raises ZeroDivisionError for ``center_freq == 0`` exactly like the reference does (cwg.py:31).
    """
    samples_per_cycle = sampling_frequency / center_freq  # cwg.py:31
    cycles = int(num_samples / samples_per_cycle)  # cwg.py:32
    if num_samples == 1:
        return 0.0  # np.linspace(0, c, 1) == [0.0]
    return cycles / (num_samples - 1)  # np.linspace step, cwg.py:33


def nco(num_samples: int, center_freq: float, sampling_frequency: float) -> np.ndarray:
    """complex64 mixing carrier exactly as cwg.generate_carrier_wave(complex=True, cw_scale=1) (cwg.py:31-36)."""
    samples_per_cycle = sampling_frequency / center_freq
    cycles = int(num_samples / samples_per_cycle)
    in_array = np.linspace(0, cycles, num_samples)
    return (np.exp(-1j * 2 * np.pi * in_array)).astype(np.complex64)


def carrier_wave(cw_scale, freq, sampling_frequency, num_samples, complex=False):
    """Noise-free restatement of cwg.generate_carrier_wave (cwg.py:6-44) used to build test tones.

    With ``noise_scale == 0`` the reference adds ``0 * truncnorm(...)`` (float32 zeros), so the result is the
    complex64 carrier (complex=True) or its float32 real part (complex=False).
    """
    cw = cw_scale * nco(num_samples, freq, sampling_frequency)
    cw = cw + np.zeros(num_samples, dtype=np.float32)  # cwg.py:39-42 with scale 0
    return cw if complex else np.real(cw)


def out_len(n_samples: int, n_taps: int, decimation: int) -> int:
    """Length of run()'s result: ceil((|N - T| + 1) / D)  (scipy 'valid' swaps operands when N < T)."""
    full = abs(n_samples - n_taps) + 1
    return -(-full // decimation)


# ----------------------------------------------------------------------------------------------------------
# Faithful restatement (same library calls, same dtypes, same order) -- the CPU baseline that is timed.
# ----------------------------------------------------------------------------------------------------------
def ddc_reference(
    input_data: np.ndarray,
    center_freq: float,
    taps: np.ndarray,
    decimation: int,
    sampling_frequency: float,
    faithful_noise: bool = False,
) -> np.ndarray:
    """DigitalDownConverter.run restated step by step (ddc.py:121-162).

    `faithful_noise=True` additionally draws the N truncated-normal samples that the reference draws and then
    multiplies by 0.0 (cwg.py:39,68-70); they do not change the result, only the run time (this is what the
    timed "reference arm" uses so that the CPU baseline is not flattered or handicapped).
    """
    from scipy import signal

    if len(input_data) == 0:  # ddc.py:137-138
        raise ValueError(f"Too few samples in input data. Received {len(input_data)}")
    n = int(np.size(input_data))
    cw = nco(n, center_freq, sampling_frequency)  # ddc.py:145-152 -> cwg.py:31-36
    if faithful_noise:
        import scipy.stats

        noise = 0.0 * scipy.stats.truncnorm.rvs(-2.0, 2.0, loc=0.0, scale=0.5, size=n).astype(np.float32)
        cw = cw + noise  # cwg.py:39-42
    mix = input_data * cw  # ddc.py:66
    filtered = signal.convolve(mix, taps, mode="valid") / sum(taps)  # ddc.py:98
    return filtered[0::decimation]  # ddc.py:119


# ----------------------------------------------------------------------------------------------------------
# Windowed float64 form: same arithmetic, O(window) memory -- used to check 2^28-sample runs on slices.
# ----------------------------------------------------------------------------------------------------------
def ddc_windowed_f64(
    x: np.ndarray,
    m_start: int,
    m_count: int,
    step: float,
    taps: np.ndarray,
    decimation: int,
    sample_offset: int = 0,
    x_base: int = 0,
) -> np.ndarray:
    """Outputs m_start .. m_start+m_count-1 of the reference for an N >= T call.

        cw[n]  = complex64(exp(-j 2 pi (n * step)))            cwg.py:33,36
        mix[n] = complex64(float32(x[n]) * cw[n])              ddc.py:66
        y[m]   = (1/sum(taps)) * sum_i taps[i] * mix[m D + T-1 - i]   ddc.py:98,119  (complex128)

    `x` holds samples x_base .. x_base+len(x)-1 of the stream (so a slice of a huge array can be passed);
    `sample_offset` shifts the NCO sample index (chunked operation; 0 reproduces a one-shot call).
    """
    taps = np.asarray(taps, dtype=np.float64)
    t = len(taps)
    d = int(decimation)
    n0 = m_start * d
    n1 = (m_start + m_count - 1) * d + t  # exclusive
    seg = np.asarray(x[n0 - x_base : n1 - x_base], dtype=np.float32)
    if len(seg) != n1 - n0:
        raise ValueError("window exceeds the samples provided")
    n = np.arange(n0, n1, dtype=np.float64) + float(sample_offset)
    cw = np.exp(-2j * np.pi * (n * step)).astype(np.complex64)
    mix = (seg * cw).astype(np.complex128)
    # y[m] = sum_k taps[T-1-k] * mix[m D + k]
    h = taps[::-1]
    idx = (np.arange(m_count) * d)[:, None] + np.arange(t)[None, :]
    return (mix[idx] @ h) / taps.sum()


def ddc_ideal_f64(x, m_start, m_count, step, taps, decimation, sample_offset=0):
    """All-float64 ideal (no complex64 rounding) -- used only to report how far the reference's own
    rounding sits from exact arithmetic, so tolerances can be judged."""
    taps = np.asarray(taps, dtype=np.float64)
    t, d = len(taps), int(decimation)
    n0, n1 = m_start * d, (m_start + m_count - 1) * d + t
    n = np.arange(n0, n1, dtype=np.float64) + float(sample_offset)
    mix = np.asarray(x[n0:n1], dtype=np.float64) * np.exp(-2j * np.pi * (n * step))
    idx = (np.arange(m_count) * d)[:, None] + np.arange(t)[None, :]
    return (mix[idx] @ taps[::-1]) / taps.sum()


# ----------------------------------------------------------------------------------------------------------
# Packed 10-bit digitiser samples.  The reference only has a stub (ddc.py:68-83: "Digitiser raw data 10bit
# and packed into 8bit words for transport"), so the format is DEFINED by this build and parity for the
# unpack stage is against this packer (DESIGN.md "Packed 10-bit format"):
#   sample k is a two's-complement 10-bit integer stored MSB-first at bit offset 10*k of a big-endian bit
#   stream; 4 samples occupy 5 bytes.
# ----------------------------------------------------------------------------------------------------------
def pack10(samples: np.ndarray) -> np.ndarray:
    s = np.asarray(samples)
    if s.ndim != 1 or len(s) % 4:
        raise ValueError("pack10 needs a 1-D array whose length is a multiple of 4")
    if s.size and (s.min() < -512 or s.max() > 511):
        raise ValueError("sample outside the 10-bit range")
    u = (s.astype(np.int64) & 0x3FF).reshape(-1, 4)
    word = (u[:, 0] << 30) | (u[:, 1] << 20) | (u[:, 2] << 10) | u[:, 3]  # 40 bits
    out = np.empty((len(u), 5), dtype=np.uint8)
    for b in range(5):
        out[:, b] = (word >> (32 - 8 * b)) & 0xFF
    return out.reshape(-1)


def unpack10(packed: np.ndarray) -> np.ndarray:
    p = np.asarray(packed, dtype=np.uint8)
    if p.ndim != 1 or len(p) % 5:
        raise ValueError("unpack10 needs a 1-D uint8 array whose length is a multiple of 5")
    b = p.reshape(-1, 5).astype(np.int64)
    word = (b[:, 0] << 32) | (b[:, 1] << 24) | (b[:, 2] << 16) | (b[:, 3] << 8) | b[:, 4]
    out = np.empty((len(b), 4), dtype=np.int64)
    for k in range(4):
        out[:, k] = (word >> (30 - 10 * k)) & 0x3FF
    out = np.where(out >= 512, out - 1024, out)
    return out.reshape(-1).astype(np.int16)


def taps_from_q17(numerators) -> np.ndarray:
    """The shipped tap files hold k / 2**17 printed with '%.5g' (tests/golden/make_golden.py verified that this
    reproduces feng/ddc/src/ddc_coeff_*.csv byte for byte); the reference parses that text with genfromtxt
    (ddc.py:46), so the float64 taps are float('%.5g' % (k / 2**17))."""
    return np.array([float("%.5g" % (int(k) / 131072.0)) for k in numerators], dtype=np.float64)
