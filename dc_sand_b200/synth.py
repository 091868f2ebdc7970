"""Synthetic digitiser streams for tests and benchmarks: tone + Gaussian noise, rounded and
clipped to the 10-bit range.  Seeded; stream s uses seed 1234 + s."""
from __future__ import annotations

import numpy as np

FS = 1712e6
FC = 100e6


def digitiser_stream(n: int, seed: int = 1234, fs: float = FS, f0: float = FC + 3.3e6, amp: float = 100.0,
                     sigma: float = 40.0) -> np.ndarray:
    """int16 samples in [-512, 511]."""
    rng = np.random.default_rng(seed)
    phi = rng.uniform(0, 2 * np.pi)
    t = np.arange(n, dtype=np.float64)
    x = amp * np.cos(2 * np.pi * (f0 / fs) * t + phi) + sigma * rng.standard_normal(n)
    return np.clip(np.rint(x), -512, 511).astype(np.int16)


def digitiser_stream_fast(n: int, seed: int = 1234, block: int = 1 << 20) -> np.ndarray:
    """Cheap variant for multi-gigasample benchmark inputs: one random block of `block` samples per seed, tiled with a
    per-tile circular shift so that tiles are not identical.  Same value range / spectrum class as digitiser_stream;
    used only where the content does not matter beyond being realistic (timing), never for golden vectors."""
    base = digitiser_stream(block, seed)
    reps = -(-n // block)
    out = np.empty(reps * block, dtype=np.int16)
    for r in range(reps):
        out[r * block : (r + 1) * block] = np.roll(base, 7919 * r)
    return out[:n]


def pack10(samples: np.ndarray) -> np.ndarray:
    """Pack int samples in [-512, 511] into the digitiser transport format used by this library:
    big-endian bit stream, 10 bits per sample MSB first, two's complement; 4 samples -> 5 bytes."""
    s = np.asarray(samples)
    if s.ndim != 1 or len(s) % 4:
        raise ValueError("pack10 needs a 1-D array whose length is a multiple of 4")
    if s.size and (int(s.min()) < -512 or int(s.max()) > 511):
        raise ValueError("sample outside the 10-bit range")
    u = (s.astype(np.uint32) & 0x3FF).reshape(-1, 4)
    out = np.empty((len(u), 5), dtype=np.uint8)
    out[:, 0] = u[:, 0] >> 2
    out[:, 1] = ((u[:, 0] & 0x3) << 6) | (u[:, 1] >> 4)
    out[:, 2] = ((u[:, 1] & 0xF) << 4) | (u[:, 2] >> 6)
    out[:, 3] = ((u[:, 2] & 0x3F) << 2) | (u[:, 3] >> 8)
    out[:, 4] = u[:, 3] & 0xFF
    return out.reshape(-1)
