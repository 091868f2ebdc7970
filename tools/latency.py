#!/usr/bin/env python
"""Host-API latency of DigitalDownConverter.run() (pageable NumPy in, complex128 out) at BASELINE configs[0] size and around it."""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dc_sand_b200 import DigitalDownConverter, synth, taps
ddc = DigitalDownConverter(16, 1712e6, taps.write_csv("ddc_coeff_107MHz.csv", tempfile.mkdtemp()))
threads = [int(a.split("=")[1]) for a in sys.argv[1:] if a.startswith("copy_threads=")]
if threads:
    ddc.set_option("copy_threads", threads[0])
for logn in (14, 17, 20, 22, 24, 26):
    n = 1 << logn
    x = synth.digitiser_stream_fast(n, 3, block=min(n, 1 << 20)).astype(np.float32)
    for _ in range(3): y = ddc.run(x, 100e6)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); y = ddc.run(x, 100e6); ts.append(time.perf_counter() - t0)
    print(f"N=2^{logn}: run() best {min(ts)*1e3:.3f} ms median {np.median(ts)*1e3:.3f} ms  -> {n / min(ts) / 1e9:.3f} Gsamples/s  variant {ddc.last_variant}")

# breakdown at 2^26: the C call alone (pageable input, preallocated pageable output) against the Python wrapper's extras
import ctypes as C
from dc_sand_b200 import _lib
n = 1 << 26
x = synth.digitiser_stream_fast(n, 3).astype(np.float32)
m = ddc.out_len(n)
out = np.empty(m, dtype=np.complex64); out[:] = 0
lib, h, step = _lib.load(), ddc._get_handle(), ddc.phase_step(n, 100e6)
for _ in range(2): lib.ddcb200_run_host_f32(h, x.ctypes.data, n, 1, n, step, 0, out.ctypes.data, m)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); lib.ddcb200_run_host_f32(h, x.ctypes.data, n, 1, n, step, 0, out.ctypes.data, m); ts.append(time.perf_counter() - t0)
t0 = time.perf_counter(); y = out.astype(np.complex128); t_as = time.perf_counter() - t0
t0 = time.perf_counter(); o2 = np.empty(m, dtype=np.complex64); o2[:] = 0; t_alloc = time.perf_counter() - t0
print(f"2^26 breakdown: C call best {min(ts)*1e3:.2f} ms ({n*4/min(ts)/1e9:.1f} GB/s in); astype(complex128) {t_as*1e3:.2f} ms; fresh complex64 alloc+touch {t_alloc*1e3:.2f} ms")
