"""Streaming front end of the fused DDC: a digitiser delivers an endless stream, the reference's `run()` takes one finite
array (ddc.py:121).  `DDCStream.push()` accepts the stream in arbitrary pieces; the native session (include/ddcb200.h,
ddcb200_session_*) carries the last T-D .. T-1 samples and the absolute sample index on the device, so the concatenated
outputs of any sequence of pushes equal ONE `run()` over the concatenated input.

NCO phase law: the reference's carrier advances `int(N fc / fs) / (N - 1)` cycles per sample for an N-sample call
(cwg.py:31-33).  Pass `total_samples=N` to reproduce a one-shot `run()` of N samples exactly; without it the stream uses
the true NCO step `fc / fs` (what the linspace law converges to for long arrays)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, cwg
from .ddc import DigitalDownConverter, _torch_stream


class DDCStream:
    def __init__(self, ddc: DigitalDownConverter, center_freq: float, n_streams: int = 1, max_chunk: int = 1 << 22,
                 total_samples: int | None = None, packed: bool = False) -> None:
        """`packed=True`: pushes are packed 10-bit bytes (uint8, 5 bytes per 4 samples; the format of
        DigitalDownConverter.run_packed), unpacked inside the fused kernel."""
        self.ddc = ddc
        self.n_streams = int(n_streams)
        self.max_chunk = int(max_chunk)
        self.packed = bool(packed)
        if total_samples is None:
            self.phase_step = float(center_freq) / float(ddc.sampling_frequency)
        else:
            self.phase_step = cwg.phase_step_cycles(int(total_samples), center_freq, ddc.sampling_frequency)
        s = C.c_void_p()
        opener = _lib.load().ddcb200_session_open_packed10 if self.packed else _lib.load().ddcb200_session_open
        _lib.check(opener(ddc._get_handle(), self.n_streams, self.max_chunk, self.phase_step, C.byref(s)), "ddcb200_session_open")
        self._s = s
        ddc._sessions.add(self)   # ddc.close() closes me first; ddc refuses to change taps / decimation while I am open

    # ------------------------------------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_s", None) is not None:
            _lib.load().ddcb200_session_close(self._s)
            self._s = None
            self.ddc._sessions.discard(self)

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def pending(self) -> int:
        """Samples per stream carried over to the next push."""
        return int(_lib.load().ddcb200_session_pending(self._s))

    @property
    def position(self) -> int:
        """Absolute index of the next sample to be pushed."""
        return int(_lib.load().ddcb200_session_position(self._s))

    def reset(self, first_sample_index: int = 0) -> None:
        _lib.check(_lib.load().ddcb200_session_reset(self._s, int(first_sample_index)), "ddcb200_session_reset")

    def out_len(self, n_samples: int) -> int:
        return int(_lib.load().ddcb200_session_out_len(self._s, int(n_samples)))

    # ------------------------------------------------------------------------------------------------------------
    def push(self, x: np.ndarray) -> np.ndarray:
        """Host arrays: [n] (one stream) or [streams, n] real samples (or uint8 [.., 5 n / 4] for a packed session)
        -> complex64 [m] or [streams, m]; m may be 0.  Pieces longer than `max_chunk` are cut and double-buffered inside
        the library."""
        a = np.asarray(x)
        one_d = a.ndim == 1
        a2 = a[None, :] if one_d else a
        if a2.ndim != 2 or a2.shape[0] != self.n_streams or a2.shape[1] == 0:
            raise ValueError(f"push needs [{self.n_streams}, n > 0] samples, got shape {a.shape}")
        if self.packed:
            a2 = np.ascontiguousarray(a2, dtype=np.uint8)
            if a2.shape[1] % 5:
                raise ValueError("packed pushes must be whole groups of 5 bytes (4 samples)")
            n = a2.shape[1] // 5 * 4
            fn = _lib.load().ddcb200_session_push_host_packed10
        else:
            a2 = np.ascontiguousarray(a2, dtype=np.float32)
            n = a2.shape[1]
            fn = _lib.load().ddcb200_session_push_host_f32
        m = self.out_len(n)
        out = np.empty((self.n_streams, max(m, 1)), dtype=np.complex64)
        got = C.c_int64(0)
        _lib.check(fn(self._s, a2.ctypes.data, n, a2.shape[1], out.ctypes.data, out.shape[1], C.byref(got)),
                   "ddcb200_session_push_host")
        assert got.value == m
        out = out[:, :m]
        return out[0] if one_d else out

    def push_tensor(self, x, out=None):
        """torch CUDA tensors, asynchronous on torch's current stream: float32 [n] or [streams, n] (uint8 [.., 5 n / 4] for a
        packed session), n <= max_chunk."""
        import torch

        if not x.is_cuda or x.device.index != self.ddc.device:
            raise ValueError(f"x must live on cuda:{self.ddc.device}")
        one_d = x.dim() == 1
        x2 = x.unsqueeze(0) if one_d else x
        want = torch.uint8 if self.packed else torch.float32
        if x2.dim() != 2 or x2.shape[0] != self.n_streams or x2.dtype != want or x2.stride(1) != 1:
            raise ValueError(f"push_tensor needs {want} [{self.n_streams}, n] with contiguous rows")
        if self.packed and x2.shape[1] % 5:
            raise ValueError("packed pushes must be whole groups of 5 bytes (4 samples)")
        n = x2.shape[1] // 5 * 4 if self.packed else x2.shape[1]
        m = self.out_len(n)
        if out is None:
            out = torch.empty((self.n_streams, m), dtype=torch.complex64, device=x.device)
        if not out.is_cuda or out.device != x.device:
            raise ValueError(f"out must live on cuda:{self.ddc.device} like x")
        out2 = out.unsqueeze(0) if out.dim() == 1 else out
        if out2.shape[0] != self.n_streams or out2.shape[1] < m or out2.dtype != torch.complex64 or out2.stride(1) != 1:
            raise ValueError("out must be complex64 [streams, >= m] with contiguous rows")
        got = C.c_int64(0)
        fn = _lib.load().ddcb200_session_push_packed10 if self.packed else _lib.load().ddcb200_session_push_f32
        _lib.check(fn(self._s, x2.data_ptr(), n, x2.stride(0), out2.data_ptr(), max(out2.stride(0), 1), C.byref(got),
                      _torch_stream(torch, x.device)), "ddcb200_session_push")
        res = out2[:, : got.value]
        return res[0] if one_d else res
