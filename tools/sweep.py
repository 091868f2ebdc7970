#!/usr/bin/env python
"""BASELINE configs[3]: tap/decimation sweep (taps 64-1024 x decimation 4-64), N = 2^26 float32 samples, back-to-back
device timing. Writes a markdown table (stdout) with both roofline fractions and the binding roofline per cell."""
import os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dc_sand_b200 import DigitalDownConverter, synth
from scipy import signal

HBM, FP32 = 6539.5, 74.4
n = 1 << 26
opts = [a.split("=") for a in sys.argv[1:] if "=" in a and not a.startswith("--")]       # key=value -> set_option
dmin = int(next((a.split("=")[1] for a in sys.argv[1:] if a.startswith("--dmin=")), 4))
x = torch.from_numpy(synth.digitiser_stream_fast(n, 1, block=1 << 22).astype(np.float32)).cuda().unsqueeze(0)
tmp = tempfile.mkdtemp()
print("| T | D | kernel | ms | Gsamples/s | GB/s (alg.) | % HBM (6539.5) | TFLOP/s | % FP32 (74.4) | binding roofline | % of binding |")
print("|---:|---:|---|---:|---:|---:|---:|---:|---:|---|---:|")
for t in (64, 128, 256, 512, 1024):
    for d in (4, 8, 16, 32, 64):
        if d < dmin: continue
        csv = os.path.join(tmp, f"t{t}_{d}.csv")
        np.savetxt(csv, signal.firwin(t, 0.8 / d), fmt="%.18e")
        ddc = DigitalDownConverter(d, 1712e6, csv)
        for k, v in opts: ddc.set_option(k, int(v))
        m = ddc.out_len(n)
        out = torch.empty((1, m), dtype=torch.complex64, device="cuda")
        for _ in range(3): ddc.run_tensor(x, 100e6, out=out)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): ddc.run_tensor(x, 100e6, out=out)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 20)
        by = n * (4 + 8.0 / d); fl = 4.0 * t * m
        gbs, tf = by / best / 1e6, fl / best / 1e9
        t_hbm, t_fp = by / (HBM * 1e6), fl / (FP32 * 1e9)
        bind = "HBM" if t_hbm >= t_fp else "FP32"
        print(f"| {t} | {d} | {ddc.last_variant.split('<')[0]} | {best:.4f} | {n / best / 1e6:.0f} | {gbs:.0f} | {100 * gbs / HBM:.1f} | {tf:.1f} | {100 * tf / FP32:.1f} | {bind} | {100 * max(t_hbm, t_fp) / best:.1f} |")
        ddc.close()
