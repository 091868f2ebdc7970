"""On-device spectral self-check: the step after the DDC in an F-engine is a channeliser, and the
reference's acceptance tests (feng/ddc/testing/test_ddc.py:61-332) judge the DDC by the power spectrum of its last
2^15 outputs.  These helpers do that check where the data already is -- an FFT of the tail of a device tensor -- and hand
back only bin indices / a level, so a deployment can verify a tone without copying baseband to the host.
torch.fft is used as a library here: this is a checker, not part of the DDC path."""
from __future__ import annotations


def power_spectrum(y, fft_length: int = 2 ** 15):
    """|FFT|^2 of the last `fft_length` outputs of a complex64 device tensor ([M] or [streams, M])."""
    import torch

    if y.shape[-1] < fft_length:
        raise ValueError(f"need at least {fft_length} outputs, got {y.shape[-1]}")
    return torch.fft.fft(y[..., -fft_length:].to(torch.complex128), dim=-1).abs() ** 2


def spectrum_bins_above(y, fft_length: int = 2 ** 15, threshold: float = 1e5):
    """Bins whose power exceeds `threshold`, as the reference's tests compute them (np.where(P > 1e5), test_ddc.py:73)."""
    p = power_spectrum(y, fft_length)
    idx = (p > threshold).nonzero(as_tuple=False)
    if p.dim() == 1:
        return [int(i) for i in idx[:, 0].cpu()]
    return [[int(b) for s, b in idx.cpu().tolist() if s == k] for k in range(p.shape[0])]


def peak_power_db(y, fft_length: int = 2 ** 15) -> float:
    """10 log10 of the strongest bin."""
    import torch

    return float(10.0 * torch.log10(power_spectrum(y, fft_length).max()))
