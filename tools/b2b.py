#!/usr/bin/env python
"""Back-to-back kernel timing (no host sync between launches): python tools/b2b.py [--iters 200] key=value ..."""
import os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dc_sand_b200 import DigitalDownConverter, synth, taps
iters = 200
opts = []
for a in sys.argv[1:]:
    if a.startswith("--iters="): iters = int(a.split("=")[1])
    else: opts.append(a)
ddc = DigitalDownConverter(16, 1712e6, taps.write_csv("ddc_coeff_107MHz.csv", tempfile.mkdtemp()))
for kv in opts:
    k, v = kv.split("="); ddc.set_option(k, int(v))
n = 1 << 28
x = torch.from_numpy(synth.digitiser_stream_fast(n, 1, block=1 << 22).astype(np.float32)).cuda().unsqueeze(0)
out = torch.empty((1, ddc.out_len(n)), dtype=torch.complex64, device="cuda")
for _ in range(5): ddc.run_tensor(x, 100e6, out=out)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ddc.run_tensor(x, 100e6, out=out)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / iters)
b = n * 4.5
print(f"{' '.join(opts):40s} {ddc.last_variant}: {best:.4f} ms  {n / best / 1e6:.1f} Gsamples/s  {b / best / 1e6:.0f} GB/s ({b / best / 1e6 / 65.395:.1f}% HBM)  {4*256*ddc.out_len(n)/best/1e9/74.4*100:.1f}% FP32")
