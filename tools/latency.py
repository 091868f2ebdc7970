#!/usr/bin/env python
"""Host-API latency of DigitalDownConverter.run() (pageable NumPy in, complex128 out) at BASELINE configs[0] size and around it."""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dc_sand_b200 import DigitalDownConverter, synth, taps
ddc = DigitalDownConverter(16, 1712e6, taps.write_csv("ddc_coeff_107MHz.csv", tempfile.mkdtemp()))
for logn in (14, 17, 20, 22, 24):
    n = 1 << logn
    x = synth.digitiser_stream_fast(n, 3, block=min(n, 1 << 20)).astype(np.float32)
    for _ in range(3): y = ddc.run(x, 100e6)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); y = ddc.run(x, 100e6); ts.append(time.perf_counter() - t0)
    print(f"N=2^{logn}: run() best {min(ts)*1e3:.3f} ms median {np.median(ts)*1e3:.3f} ms  -> {n / min(ts) / 1e9:.3f} Gsamples/s  variant {ddc.last_variant}")
