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


# ---------------------------------------------------------------------------------------------------------------------
# Device-side generator: the same carrier wave produced directly in HBM by libddcb200.so
# (ddcb200_cwg), for inputs that should never be staged through the host.
# ---------------------------------------------------------------------------------------------------------------------
_gen_handles: dict = {}


def _generator_handle(device: int):
    """A minimal native handle (one unit tap) per device: ddcb200_cwg only needs its device and stream."""
    import ctypes as C

    from . import _lib

    if device not in _gen_handles:
        h = C.c_void_p()
        one = (C.c_double * 1)(1.0)
        _lib.check(_lib.load().ddcb200_create(C.byref(h), int(device), one, 1, 1), "ddcb200_create")
        _gen_handles[device] = h
    return _gen_handles[device]


def generate_carrier_wave_gpu(cw_scale: float, freq: float, sampling_frequency: float, num_samples: int, noise_scale: float,
                              complex: bool, seed: int = 0, device: int = 0, n_streams: int | None = None,
                              sample_offset: int = 0, digitise: bool = False, phase0_cycles: float = 0.0, out=None,
                              total_samples: int | None = None):
    """`generate_carrier_wave` on the GPU: returns a torch CUDA tensor, float32 (complex=False) or complex64, of shape
    [num_samples] (or [n_streams, num_samples]; stream s draws its noise from key (seed, s)).

    The tone follows the reference's phase law exactly (cwg.py:31-36); the noise is the reference's truncated normal
    (sigma 0.5 on [-1, 1], cwg.py:47-70) from a seeded Philox counter RNG -- the reference's own draw is unseeded, so only
    its statistics can be matched.  digitise=True switches to the digitiser model of SURVEY 8d instead: real part plus
    noise_scale * N(0, 1), rounded and clipped to the 10-bit range."""
    import torch

    from . import _lib
    from .ddc import _torch_stream

    dev = torch.device("cuda", int(device))
    s = 1 if n_streams is None else int(n_streams)
    if out is None:
        out = torch.empty((s, int(num_samples)), dtype=torch.complex64 if complex else torch.float32, device=dev)
    elif out.dim() != 2 or out.shape != (s, int(num_samples)) or out.stride(1) != 1:
        raise ValueError("out must be [n_streams, num_samples] with contiguous rows")
    # a piece of a longer wave: phase law of the whole (total_samples), starting at sample_offset
    step = phase_step_cycles(int(num_samples if total_samples is None else total_samples), freq, sampling_frequency)
    mode = 0 if noise_scale == 0 else (2 if digitise else 1)
    _lib.check(
        _lib.load().ddcb200_cwg(_generator_handle(int(device)), out.data_ptr(), int(num_samples), s, out.stride(0), int(bool(complex)),
                                float(cw_scale), step, float(phase0_cycles), int(sample_offset), mode, float(noise_scale),
                                int(seed) & 0xFFFFFFFFFFFFFFFF, _torch_stream(torch, dev)),
        "ddcb200_cwg",
    )
    return out[0] if n_streams is None else out


def pack10_gpu(x, out=None):
    """Digitiser transport format built in HBM (ddcb200_pack10): float32 CUDA tensor [n] or [streams, n] (n % 4 == 0,
    contiguous rows) -> uint8 [.., 5 n / 4]; samples are rounded to nearest and clipped to [-512, 511].  The inverse of
    DigitalDownConverter._decode_8bit_to_10bit_to_float_data, for packed test vectors that never touch the host."""
    import torch

    from . import _lib
    from .ddc import _torch_stream

    one_d = x.dim() == 1
    x2 = x.unsqueeze(0) if one_d else x
    if not x2.is_cuda or x2.dim() != 2 or x2.dtype != torch.float32 or x2.stride(1) != 1 or x2.shape[1] % 4:
        raise ValueError("pack10_gpu needs float32 CUDA rows with contiguous samples, n % 4 == 0")
    s, n = x2.shape
    if out is None:
        out = torch.empty((s, n // 4 * 5), dtype=torch.uint8, device=x.device)
    out2 = out.unsqueeze(0) if out.dim() == 1 else out
    if out2.shape != (s, n // 4 * 5) or out2.dtype != torch.uint8 or out2.stride(1) != 1:
        raise ValueError("out must be uint8 [streams, 5 n / 4] with contiguous rows")
    dev = x.device.index
    _lib.check(
        _lib.load().ddcb200_pack10(_generator_handle(dev), x2.data_ptr(), n, s, x2.stride(0), out2.data_ptr(), out2.stride(0),
                                   _torch_stream(torch, x.device)),
        "ddcb200_pack10",
    )
    return out2[0] if one_d else out2
