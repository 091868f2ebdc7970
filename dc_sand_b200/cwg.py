"""Carrier-wave generator with the reference's name, signature and phase law (feng/ddc/src/cwg.py:6-70).

This is the *test-vector source* of the reference (its tests build their inputs with it) and it also defines the
NCO phase law that `DigitalDownConverter.run` reproduces on the GPU.  It runs on the host in NumPy, as in the
reference; the NCO used on the hot path is generated inside the CUDA kernel, not here.
"""
from __future__ import annotations

import numpy as np


def phase_step_cycles(num_samples: int, freq: float, sampling_frequency: float) -> float:
    """NCO cycles per sample implied by cwg.py:31-33: linspace(0, int(N / (fs / f)), N) has step cycles/(N-1)."""
    samples_per_cycle = sampling_frequency / freq  # ZeroDivisionError for freq == 0, like the reference
    cycles = int(num_samples / samples_per_cycle)
    if num_samples <= 1:
        return 0.0
    return cycles / (num_samples - 1)


def generate_carrier_wave(
    cw_scale: float, freq: float, sampling_frequency: int, num_samples: int, noise_scale: float, complex: bool
) -> np.ndarray:
    """Generate a carrier wave vector (same arguments and dtypes as feng/ddc/src/cwg.py:6-44).

    complex=True  -> complex64 ``cw_scale * exp(-j 2 pi n step)`` (+ float32 noise on the real part)
    complex=False -> float32 real part of the same.
    """
    samples_per_cycle = sampling_frequency / freq
    cycles = int(num_samples / samples_per_cycle)
    in_array = np.linspace(0, cycles, num_samples)
    carrier_wave_complex = cw_scale * (np.exp(-1j * 2 * np.pi * in_array)).astype(np.complex64)
    additive_white_gaussian_noise = _generate_noise(noise_scale, len(carrier_wave_complex))
    if complex is True:
        return carrier_wave_complex + additive_white_gaussian_noise
    return np.real(carrier_wave_complex + additive_white_gaussian_noise)


def _generate_noise(scale: float, array_length: int, rng: np.random.Generator | None = None) -> np.ndarray:
    """Truncated-normal noise on [-1, 1], sigma 0.5, float32 (cwg.py:47-70).

    Unlike the reference, nothing is drawn when ``scale == 0`` (the reference draws N samples and multiplies them
    by zero, which is 60-75 % of its run time), and an optional seeded generator makes vectors reproducible.
    """
    if scale == 0:
        return np.zeros(array_length, dtype=np.float32)
    rng = np.random.default_rng() if rng is None else rng
    sigma, lo, hi = 0.5, -1.0, 1.0
    out = np.empty(array_length, dtype=np.float64)
    filled = 0
    while filled < array_length:  # rejection sampling: 95.4 % acceptance
        draw = rng.normal(0.0, sigma, size=int((array_length - filled) * 1.1) + 16)
        draw = draw[(draw >= lo) & (draw <= hi)]
        take = min(len(draw), array_length - filled)
        out[filled : filled + take] = draw[:take]
        filled += take
    return scale * out.astype(np.float32)
