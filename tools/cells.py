#!/usr/bin/env python
"""Back-to-back timing of several (taps, decimation, variant, log2 n) cells in ONE process, each checked against the float64
windowed oracle:
    python tools/cells.py 64,16,11,26 64,16,0,26 ...
"""
import os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dc_sand_b200 import DigitalDownConverter, synth, taps
from oracle import ddc_oracle as orc
from scipy import signal

HBM = 6539.5
tmp = tempfile.mkdtemp()
xs = {}
for spec in sys.argv[1:]:
    T, D, var, logn = (int(v) for v in spec.split(","))
    n = 1 << logn
    if n not in xs:
        xs[n] = synth.digitiser_stream_fast(n, 1, block=min(n, 1 << 22)).astype(np.float32)
    xh = xs[n]
    x = torch.from_numpy(xh).cuda().unsqueeze(0)
    if T == 256:
        csv = taps.write_csv("ddc_coeff_107MHz.csv", tmp)
    else:
        csv = os.path.join(tmp, f"t{T}_{D}.csv"); np.savetxt(csv, signal.firwin(T, 0.8 / D), fmt="%.18e")
    ddc = DigitalDownConverter(D, 1712e6, csv)
    ddc.set_option("variant", var)
    m = ddc.out_len(n)
    out = torch.empty((1, m), dtype=torch.complex64, device="cuda")
    for _ in range(3): ddc.run_tensor(x, 100e6, out=out)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): ddc.run_tensor(x, 100e6, out=out)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 50)
    step = orc.phase_step_cycles(n, 100e6, 1712e6)
    worst = 0.0
    yo = out[0].cpu().numpy()
    scale = np.abs(yo[:4096]).max()
    for m0 in (0, 12345, m // 2 + 7, m - 600):
        ref = orc.ddc_windowed_f64(xh, m0, 512, step, ddc.ddc_filter_coeffs, D)
        worst = max(worst, np.abs(yo[m0:m0 + 512] - ref).max() / scale)
    by = n * (4 + 8.0 / D)
    print(f"T={T} D={D} variant={var} n=2^{logn}: {ddc.last_variant}: {best:.4f} ms  {by / best / 1e6:.0f} GB/s ({100 * by / best / 1e6 / HBM:.1f}% HBM)  "
          f"{4.0 * T * m / best / 1e9:.1f} TFLOP/s direct-form  err {worst:.2e}", flush=True)
    ddc.close()
