"""Digital Down Conversion for the F-engine -- B200 drop-in for feng/ddc/src/ddc.py.

Same class name, constructor arguments, attributes and `run(input_data, center_freq)` contract as the reference
(`/root/reference/feng/ddc/src/ddc.py:10-188`), but `run` executes NCO mixing, FIR filtering and decimation in
one fused sm_100a CUDA kernel through the C ABI of `libddcb200.so` (include/ddcb200.h).  No NumPy/SciPy
arithmetic happens on the hot path and there is no CPU fallback: without the library `run` raises.

Extensions that the reference does not have (all optional keyword arguments or extra methods):
  * `device=` constructor argument, `run_batch` (many streams per call), `run_packed` / `run_batch_packed`
    (packed 10-bit digitiser input, the reference's stub `_decode_8bit_to_10bit_to_float_data`),
    `run_tensor` (torch CUDA tensors in and out, asynchronous), `sample_offset=`/`total_samples=` for chunked
    operation with a continuous NCO phase.
"""
from __future__ import annotations

import ctypes as C
import logging
import weakref

import numpy as np
from numpy import genfromtxt

from . import _lib, cwg

_log = logging.getLogger(__name__)


def _torch_stream(torch, device):
    """cudaStream_t of torch's current stream.  torch's default stream is the legacy NULL stream, but NULL means "the
    handle's own stream" in the C ABI, so it is passed as cudaStreamLegacy (0x1)."""
    st = torch.cuda.current_stream(device).cuda_stream
    return st if st else 1


class DigitalDownConverter:
    """Digital Down Conversion (reference: ddc.py:10)."""

    def __init__(self, decimation_factor: int, sampling_frequency: int, ddc_coeff_filename: str, device: int = 0) -> None:
        """Same parameters as the reference (ddc.py:13-31); `device` selects the CUDA device of this object."""
        self.decimation_factor = decimation_factor
        self._import_ddc_filter_coeffs(filename=ddc_coeff_filename)
        self.sampling_frequency = sampling_frequency
        self.device = int(device)
        self._handle = None
        self._handle_key = None
        self._sessions = weakref.WeakSet()   # open DDCStream sessions: they hold the native handle and its (T, D) geometry

    # ------------------------------------------------------------------------------------------------ taps
    def _import_ddc_filter_coeffs(self, filename: str = "ddc_filter_coeffs_107.csv"):
        """Import the FIR coefficients from a one-value-per-line CSV (ddc.py:33-48)."""
        _log.info("Importing coefficients from %s", filename)
        ddc_coeffs = genfromtxt(filename, delimiter=",")
        _log.info("Imported %d coefficients", len(ddc_coeffs))
        self.ddc_filter_coeffs = ddc_coeffs

    # ------------------------------------------------------------------------------------------------ handle
    def _get_handle(self):
        """Create (or refresh, if the public attributes were changed) the native handle."""
        lib = _lib.load()
        if int(self.decimation_factor) != self.decimation_factor or int(self.decimation_factor) <= 0:
            raise ValueError(f"decimation_factor must be a positive integer, got {self.decimation_factor!r}")
        taps = np.ascontiguousarray(self.ddc_filter_coeffs, dtype=np.float64).reshape(-1)
        key = (int(self.decimation_factor), taps.tobytes(), self.device)
        if self._handle is not None and key == self._handle_key:
            return self._handle
        if self._handle is not None and any(s._s is not None for s in self._sessions):
            # a session's carry, pitch and output capacity were sized for the old taps / decimation
            raise RuntimeError("decimation_factor / ddc_filter_coeffs changed while a DDCStream of this object is open; "
                               "close the stream first")
        if self._handle is None:
            h = C.c_void_p()
            _lib.check(
                lib.ddcb200_create(C.byref(h), self.device, taps.ctypes.data_as(C.POINTER(C.c_double)), len(taps),
                                   int(self.decimation_factor)),
                "ddcb200_create",
            )
            self._handle = h
        else:
            _lib.check(lib.ddcb200_set_taps(self._handle, taps.ctypes.data_as(C.POINTER(C.c_double)), len(taps)))
            _lib.check(lib.ddcb200_set_decimation(self._handle, int(self.decimation_factor)))
        self._handle_key = key
        return self._handle

    def close(self):
        for s in list(getattr(self, "_sessions", ())):   # sessions dereference the handle: they go first
            s.close()
        if getattr(self, "_handle", None) is not None:
            _lib.load().ddcb200_destroy(self._handle)
            self._handle = None
            self._handle_key = None

    def __del__(self):  # pragma: no cover - interpreter shutdown order
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------ helpers
    def out_len(self, num_samples: int) -> int:
        """Length of run()'s result for `num_samples` inputs: ceil((|N - T| + 1) / D)."""
        return int(_lib.load().ddcb200_out_len(int(num_samples), len(self.ddc_filter_coeffs), int(self.decimation_factor)))

    def phase_step(self, num_samples: int, center_freq: float) -> float:
        """NCO cycles per sample the reference would use for a call of `num_samples` samples (cwg.py:31-33)."""
        return cwg.phase_step_cycles(num_samples, center_freq, self.sampling_frequency)

    @property
    def launch_count(self) -> int:
        return int(_lib.load().ddcb200_launch_count(self._get_handle()))

    @property
    def last_variant(self) -> str:
        return _lib.load().ddcb200_last_variant(self._get_handle()).decode()

    def set_option(self, key: str, value: int) -> None:
        _lib.check(_lib.load().ddcb200_set_option(self._get_handle(), key.encode(), int(value)))

    # ------------------------------------------------------------------------------------------------ run
    def run(self, input_data: np.ndarray, center_freq: float, sample_offset: int = 0,
            total_samples: int | None = None) -> np.ndarray:
        """Digital down-conversion of one real 1-D stream (ddc.py:121-188).

        Returns the translated, filtered and decimated complex baseband, dtype complex128, length
        ``floor((N - T) / D) + 1`` -- element m is aligned to input window ``[m D, m D + T)`` exactly like
        ``convolve(x * nco, taps, "valid")[0::D] / sum(taps)`` in the reference.

        `total_samples` / `sample_offset` (extensions): when a long stream is processed in chunks, pass the length
        of the whole stream and the index of this chunk's first sample so that the NCO phase law
        (``int(N fc / fs) / (N - 1)`` cycles per sample, cwg.py:31-33) is that of the one-shot call.
        """
        # Sanity check the input data (ddc.py:137-138).
        if len(input_data) == 0:
            raise ValueError(f"Too few samples in input data. Received {len(input_data)}")
        x = np.asarray(input_data)
        if x.ndim != 1:
            # the reference fails in _mix with a broadcasting ValueError for anything but 1-D input
            raise ValueError(f"operands could not be broadcast together: input_data must be 1-D, got shape {x.shape}")
        if np.iscomplexobj(x):
            # The reference multiplies whatever it is given by the carrier (ddc.py:66), so complex input works there; the
            # operator is linear, so here the real and imaginary parts go through the kernel as two streams of one batch.
            parts = np.stack([x.real, x.imag]).astype(np.float32)
            yb = self.run_batch(parts, center_freq, sample_offset=sample_offset, total_samples=total_samples) \
                if parts.shape[1] >= len(self.ddc_filter_coeffs) else \
                np.stack([self.run(p, center_freq, sample_offset, total_samples).astype(np.complex64) for p in parts])
            return yb[0].astype(np.complex128) + 1j * yb[1].astype(np.complex128)
        # Any real dtype the reference accepts (int8 .. float64) is converted to float32, the device's sample type: exact for
        # digitiser data (<= 16-bit integers); float64 input is rounded to float32 where the reference would keep it (6e-8
        # relative, far inside the parity tolerance)
        x = np.ascontiguousarray(x, dtype=np.float32)
        n = x.shape[0]
        step = self.phase_step(n if total_samples is None else int(total_samples), center_freq)
        h = self._get_handle()
        lib = _lib.load()
        m = self.out_len(n)
        # complex128 like the reference; widened from the device's complex64 inside the library (on host threads, chunk by
        # chunk under the transfers) rather than by a second pass here
        out = np.empty(m, dtype=np.complex128)
        _lib.check(
            lib.ddcb200_run_host_f32_c128(h, x.ctypes.data, n, step, int(sample_offset), out.ctypes.data),
            "ddcb200_run_host_f32_c128",
        )
        return out

    def run_batch(self, input_data: np.ndarray, center_freq: float, sample_offset: int = 0,
                  total_samples: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
        """`run` for many independent streams: input [streams, N] real -> output [streams, M] complex64."""
        x = np.asarray(input_data)
        if x.ndim != 2 or x.shape[0] == 0 or x.shape[1] == 0:
            raise ValueError(f"run_batch needs a non-empty [streams, N] array, got shape {x.shape}")
        x = np.ascontiguousarray(x, dtype=np.float32)
        s, n = x.shape
        if n < len(self.ddc_filter_coeffs):
            raise ValueError(f"Too few samples in input data. Received {n} < {len(self.ddc_filter_coeffs)} taps")
        step = self.phase_step(n if total_samples is None else int(total_samples), center_freq)
        m = self.out_len(n)
        if out is None:
            out = np.empty((s, m), dtype=np.complex64)
        elif out.shape != (s, m) or out.dtype != np.complex64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous complex64 array of shape [streams, M]")
        _lib.check(
            _lib.load().ddcb200_run_host_f32(self._get_handle(), x.ctypes.data, n, s, n, step, int(sample_offset),
                                             out.ctypes.data, m),
            "ddcb200_run_host_f32",
        )
        return out

    # ---- the reference's stage methods (run() itself is fused; these keep the stage API on the GPU) ----------------
    def _stage_call(self, fn_name, arrays, n_out, make_args):
        """Copy host arrays to the device, run one stage entry point of the C ABI, return the complex64 result."""
        import torch

        dev = torch.device("cuda", self.device)
        d = [torch.from_numpy(a).to(dev) for a in arrays]
        out = torch.empty(max(n_out, 1), dtype=torch.complex64, device=dev)
        fn = getattr(_lib.load(), fn_name)
        ptrs = [t.data_ptr() for t in d]
        _lib.check(fn(self._get_handle(), *ptrs, *make_args(out), _torch_stream(torch, dev)), fn_name)
        return out[:n_out].cpu().numpy()

    def _mix(self, mixing_carrier_wave: np.ndarray, input_data: np.ndarray) -> np.ndarray:
        """Multiply mixing CW with input data (ddc.py:50-66): float32[N] x complex64[N] -> complex64[N]."""
        cw = np.ascontiguousarray(mixing_carrier_wave, dtype=np.complex64)
        x = np.ascontiguousarray(input_data, dtype=np.float32)
        if x.ndim != 1 or cw.shape != x.shape:
            raise ValueError(f"operands could not be broadcast together with shapes {x.shape} {cw.shape}")
        if x.size == 0:
            return np.empty(0, dtype=np.complex64)
        return self._stage_call("ddcb200_mix_f32", [x, cw], x.size, lambda out: (out.data_ptr(), x.size))

    def _bandpass_fir_filter(self, input_data: np.ndarray) -> np.ndarray:
        """Full-rate "valid" FIR divided by sum(taps) (ddc.py:85-100): complex64[N] -> complex128[N - T + 1].

        The device computes in float32 (complex64); the result is widened to the reference's complex128."""
        z = np.ascontiguousarray(input_data, dtype=np.complex64)
        n_taps = len(self.ddc_filter_coeffs)
        if z.ndim != 1 or z.size < n_taps:
            raise ValueError(f"Too few samples in input data. Received {z.size} < {n_taps} taps")
        n_out = z.size - n_taps + 1
        y = self._stage_call("ddcb200_fir_c64", [z], n_out, lambda out: (z.size, out.data_ptr()))
        return y.astype(np.complex128)

    def _decimate(self, input_data: np.ndarray, decimate_offset: int = 0) -> np.ndarray:
        """Keep every decimation_factor-th sample starting at decimate_offset (ddc.py:102-119)."""
        src = np.asarray(input_data)
        z = np.ascontiguousarray(src, dtype=np.complex64)
        off = int(decimate_offset)
        if z.ndim != 1 or off < 0:
            raise ValueError("input_data must be 1-D and decimate_offset >= 0")
        d = int(self.decimation_factor)
        n_out = max(0, -(-(z.size - off) // d))
        if n_out == 0:
            return np.empty(0, dtype=src.dtype if np.iscomplexobj(src) else np.complex64)
        y = self._stage_call("ddcb200_decimate_c64", [z], n_out, lambda out: (z.size, off, out.data_ptr()))
        return y.astype(np.complex128) if src.dtype == np.complex128 else y

    # ---- packed 10-bit input (reference stub: ddc.py:68-83) ---------------------------------------------------
    def _decode_8bit_to_10bit_to_float_data(self, data_8bit: np.ndarray) -> np.ndarray:
        """Convert 8-bit-packed 10-bit digitiser samples to float32 on the GPU (the reference only has `pass`).

        Format: big-endian bit stream, 10-bit two's-complement samples MSB first, 4 samples per 5 bytes.
        """
        import torch

        p = np.ascontiguousarray(data_8bit, dtype=np.uint8).reshape(-1)
        if len(p) % 5:
            raise ValueError("packed data length must be a multiple of 5 bytes")
        n = len(p) // 5 * 4
        dev = torch.device("cuda", self.device)
        d_in = torch.from_numpy(p).to(dev)
        d_out = torch.empty(n, dtype=torch.float32, device=dev)
        st = _torch_stream(torch, dev)
        _lib.check(_lib.load().ddcb200_unpack10(self._get_handle(), d_in.data_ptr(), n, None, d_out.data_ptr(), st))
        return d_out.cpu().numpy()

    def run_packed(self, packed: np.ndarray, center_freq: float, sample_offset: int = 0,
                   total_samples: int | None = None) -> np.ndarray:
        """`run` on packed 10-bit input with the unpack fused into the kernel's load path. Returns complex128."""
        p = np.ascontiguousarray(packed, dtype=np.uint8)
        if p.ndim != 1 or len(p) == 0 or len(p) % 5:
            raise ValueError("packed input must be a non-empty 1-D uint8 array whose length is a multiple of 5")
        n = len(p) // 5 * 4
        if n < len(self.ddc_filter_coeffs):
            raise ValueError(f"Too few samples in input data. Received {n} < {len(self.ddc_filter_coeffs)} taps")
        step = self.phase_step(n if total_samples is None else int(total_samples), center_freq)
        m = self.out_len(n)
        out = np.empty(m, dtype=np.complex64)
        _lib.check(
            _lib.load().ddcb200_run_host_packed10(self._get_handle(), p.ctypes.data, n, 1, len(p), step,
                                                  int(sample_offset), out.ctypes.data, m),
            "ddcb200_run_host_packed10",
        )
        return out.astype(np.complex128)

    def run_batch_packed(self, packed: np.ndarray, center_freq: float, sample_offset: int = 0,
                         total_samples: int | None = None) -> np.ndarray:
        p = np.ascontiguousarray(packed, dtype=np.uint8)
        if p.ndim != 2 or p.shape[1] == 0 or p.shape[1] % 5:
            raise ValueError("packed input must be [streams, 5*k] uint8")
        s, nb = p.shape
        n = nb // 5 * 4
        step = self.phase_step(n if total_samples is None else int(total_samples), center_freq)
        m = self.out_len(n)
        out = np.empty((s, m), dtype=np.complex64)
        _lib.check(
            _lib.load().ddcb200_run_host_packed10(self._get_handle(), p.ctypes.data, n, s, nb, step, int(sample_offset),
                                                  out.ctypes.data, m),
            "ddcb200_run_host_packed10",
        )
        return out

    # ---- device-resident tensors ----------------------------------------------------------------------------------
    def run_tensor(self, x, center_freq: float, out=None, sample_offset: int = 0, total_samples: int | None = None,
                   packed: bool = False):
        """Asynchronous run on torch CUDA tensors (launched on torch's current stream).

        x: float32 [N] or [streams, N] (or uint8 [.., 5*N/4] when packed=True), contiguous rows, on this object's
        device.  Returns complex64 [M] or [streams, M].  torch is only the memory/stream plumbing here.
        """
        import torch

        if not x.is_cuda or x.device.index != self.device:
            raise ValueError(f"x must live on cuda:{self.device}")
        squeeze = x.dim() == 1
        x2 = x.unsqueeze(0) if squeeze else x
        if x2.dim() != 2 or x2.stride(1) != 1:
            raise ValueError("x must be [N] or [streams, N] with contiguous rows")
        s = x2.shape[0]
        if packed:
            if x2.dtype != torch.uint8 or x2.shape[1] % 5:
                raise ValueError("packed input must be uint8 with 5*k bytes per stream")
            n = x2.shape[1] // 5 * 4
        else:
            if x2.dtype != torch.float32:
                raise ValueError("x must be float32")
            n = x2.shape[1]
        if n < len(self.ddc_filter_coeffs):
            raise ValueError(f"Too few samples in input data. Received {n} < {len(self.ddc_filter_coeffs)} taps")
        step = self.phase_step(n if total_samples is None else int(total_samples), center_freq)
        m = self.out_len(n)
        if out is None:
            out = torch.empty((s, m), dtype=torch.complex64, device=x.device)
        if not out.is_cuda or out.device != x.device:
            raise ValueError(f"out must live on cuda:{self.device} like x")   # a host pointer would fault inside the kernel
        out2 = out.unsqueeze(0) if out.dim() == 1 else out
        if out2.shape != (s, m) or out2.dtype != torch.complex64 or out2.stride(1) != 1:
            raise ValueError("out must be complex64 [streams, M] with contiguous rows")
        st = _torch_stream(torch, x.device)
        lib = _lib.load()
        fn = lib.ddcb200_run_packed10 if packed else lib.ddcb200_run_f32
        _lib.check(
            fn(self._get_handle(), x2.data_ptr(), n, s, x2.stride(0), step, int(sample_offset), out2.data_ptr(),
               out2.stride(0), st),
            "ddcb200_run",
        )
        return out2[0] if squeeze else out2
