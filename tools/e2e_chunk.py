#!/usr/bin/env python
"""End-to-end (pinned host buffers -> ddcb200_run_host_f32 -> pinned host) time of the headline config for several time-chunk
sizes of the host path (option chunk_samples)."""
import os, sys, tempfile, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dc_sand_b200 import DigitalDownConverter, _lib, synth, taps
ddc = DigitalDownConverter(16, 1712e6, taps.write_csv("ddc_coeff_107MHz.csv", tempfile.mkdtemp()))
n = 1 << 28
x = torch.from_numpy(synth.digitiser_stream_fast(n, 1, block=1 << 22).astype(np.float32)).pin_memory()
m = ddc.out_len(n)
out = torch.empty(m, dtype=torch.complex64).pin_memory()
lib, h, step = _lib.load(), ddc._get_handle(), ddc.phase_step(n, 100e6)
for logc in (24, 23, 22, 21, 20):
    ddc.set_option("chunk_samples", 1 << logc)
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        _lib.check(lib.ddcb200_run_host_f32(h, x.data_ptr(), n, 1, n, step, 0, out.data_ptr(), m))
        ts.append(time.perf_counter() - t0)
    print(f"chunk 2^{logc}: best {min(ts[1:])*1e3:.2f} ms -> {n/min(ts[1:])/1e9:.2f} Gsamples/s ({n*4/min(ts[1:])/1e9:.1f} GB/s H2D)")
