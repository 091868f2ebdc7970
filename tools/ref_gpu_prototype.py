#!/usr/bin/env python
"""Times the REFERENCE's own CUDA prototype of the path (feng/ddc/src/ddc_kernel.cu:11-160, `kernel_ddc`, launched as
feng/ddc/src/ddc_host_gpu.py:111-120 does: block 256, grid N/4096 - 1, four float arrays of N elements besides the input)
recompiled for sm_100a (oracle/build_ref.sh -> oracle/_ref/*.cubin), next to this repo's fused kernel on the same input:
BASELINE configs[1], 2^28 float32 samples, 256 taps, decimation 16.  The prototype is a timing baseline only -- its NCO runs
in float32 with a different phase step, so its output is checked against a float64 restatement of ITS OWN
arithmetic at the shipped size (2^14 samples), not against the NumPy reference.

    python tools/ref_gpu_prototype.py            # needs oracle/_ref/kernel_ddc_2p14.cubin and kernel_ddc_2p28.cubin
"""
import ctypes
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cuda.bindings import driver as cu  # noqa: E402

from dc_sand_b200 import DigitalDownConverter, synth, taps  # noqa: E402

FS, FC = 1712e6, 100e6
SYM = b"_Z10kernel_ddcPfS_S_fS_S_"


def ck(res):
    if res[0] != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"CUDA driver error {res[0]}")
    return res[1] if len(res) == 2 else res[1:]


def load(logn):
    path = os.path.join(ROOT, "oracle", "_ref", f"kernel_ddc_2p{logn}.cubin")
    mod = ck(cu.cuModuleLoadData(open(path, "rb").read()))
    return mod, ck(cu.cuModuleGetFunction(mod, SYM))


def launch(fn, n, x, coeffs, out, dbg_re, dbg_im):
    types = (ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p)
    vals = (x.data_ptr(), coeffs.data_ptr(), out.data_ptr(), FC, dbg_re.data_ptr(), dbg_im.data_ptr())
    ck(cu.cuLaunchKernel(fn, n // 4096 - 1, 1, 1, 256, 1, 1, 0, torch.cuda.current_stream().cuda_stream, (vals, types), 0))


def main():
    torch.zeros(1, device="cuda")
    tmp = tempfile.mkdtemp()
    csv = taps.write_csv("ddc_coeff_107MHz.csv", tmp)
    ddc = DigitalDownConverter(16, FS, csv)
    c = np.asarray(ddc.ddc_filter_coeffs, dtype=np.float64)
    coeffs = torch.from_numpy(c.astype(np.float32)).cuda()

    # ---- the prototype as shipped (N = 2^14) against a float64 restatement of its own arithmetic
    n = 1 << 14
    xh = synth.digitiser_stream(n, 1234).astype(np.float32)
    x = torch.from_numpy(xh).cuda()
    out = torch.zeros(2 * 256 * (n // 4096 - 1), dtype=torch.float32, device="cuda")
    d_re, d_im = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    mod14, fn14 = load(14)
    launch(fn14, n, x, coeffs, out, d_re, d_im)
    torch.cuda.synchronize()
    y = out.cpu().numpy()
    y = y[0::2] + 1j * y[1::2]
    # ddc_kernel.cu:57-62 in float32.  `cycles/N` expands to `cycles/8192*2` (the macro has no parentheses), i.e. 4 cycles / N:
    # with the `/2` of line 67 the angle is -2 n fc / Fs half-turns -- the correct NCO after all.
    cycles = np.float32(np.float32(n) / np.float32(FS / float(np.float32(FC))))
    step = np.float32(np.float32(cycles / np.float32(n // 2)) * np.float32(2))
    idx = np.arange(n)
    ang = (-(idx * np.float64(step)) / 2).astype(np.float32).astype(np.float64)            # ddc_kernel.cu:67 (float32 angle)
    mixed = xh * np.exp(1j * np.pi * ang)
    ref = np.empty(len(y), dtype=np.complex128)
    for b in range(n // 4096 - 1):
        for t in range(256):
            base = (b + 1) * 4096 + 16 * t
            ref[b * 256 + t] = np.dot(mixed[base - 255: base + 1][::-1], c.astype(np.float32))
    err = np.abs(y - ref).max() / np.abs(ref).max()
    print(f"prototype as shipped (2^14 samples, grid 3): max |out - float64 restatement of its own arithmetic| / max = {err:.2e}")
    assert err < 1e-3, err
    ck(cu.cuModuleUnload(mod14))

    # ---- the prototype with N = 2^28 against this repo's kernel, same input
    n = 1 << 28
    xh = synth.digitiser_stream_fast(n, 1234, block=1 << 22).astype(np.float32)
    x = torch.from_numpy(xh).cuda()
    mod28, fn28 = load(28)                       # 2 GiB of static __device__ arrays (mixed_data_re / _im)
    out = torch.zeros(2 * 256 * (n // 4096 - 1), dtype=torch.float32, device="cuda")
    d_re, d_im = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")

    def timed(f, iters):
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                f()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / iters)
        return best

    t_ref = timed(lambda: launch(fn28, n, x, coeffs, out, d_re, d_im), 10)
    assert bool(torch.isfinite(out).all()) and float(out.abs().max()) > 0
    ck(cu.cuModuleUnload(mod28))
    del d_re, d_im
    xo = x.unsqueeze(0)
    yo = torch.empty((1, ddc.out_len(n)), dtype=torch.complex64, device="cuda")
    t_new = timed(lambda: ddc.run_tensor(xo, FC, out=yo), 20)
    print(f"2^28 float32 samples, 256 taps, decimation 16, one B200, back-to-back launches, CUDA events:")
    print(f"  reference prototype kernel_ddc (sm_100a recompile): {t_ref:8.3f} ms  {n / t_ref / 1e6:7.1f} Gsamples/s"
          f"  (moves 20 B/sample: input + 4 float arrays of N; taps and mixed samples re-read from global / L2)")
    print(f"  this repo, {ddc.last_variant}: {t_new:8.3f} ms  {n / t_new / 1e6:7.1f} Gsamples/s")
    print(f"  ratio: {t_ref / t_new:.1f}x")


if __name__ == "__main__":
    main()
