#!/usr/bin/env python
"""Kernel-only timing of the fused DDC (device-resident input), for quick variant sweeps and as the ncu target.

    python tools/kbench.py [--taps 256] [--decim 16] [--samples 2**28] [--streams 1] [--packed] [--iters 20] [--check]
"""
import argparse
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dc_sand_b200 import DigitalDownConverter, synth, taps  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--taps", type=int, default=256)
    ap.add_argument("--decim", type=int, default=16)
    ap.add_argument("--samples", type=str, default="2**28")
    ap.add_argument("--streams", type=int, default=1)
    ap.add_argument("--packed", action="store_true")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--dbg", action="store_true", help="print the kernels' wait / total cycle counters (option dbg_counters)")
    ap.add_argument("--option", action="append", default=[], help="key=value passed to ddcb200_set_option")
    a = ap.parse_args()
    n = int(eval(a.samples))
    tmp = tempfile.mkdtemp()
    if a.taps == 256:
        csv = taps.write_csv("ddc_coeff_107MHz.csv", tmp)
    else:
        from scipy import signal

        csv = os.path.join(tmp, "t.csv")
        np.savetxt(csv, signal.firwin(a.taps, 0.8 / a.decim), fmt="%.18e")
    ddc = DigitalDownConverter(a.decim, 1712e6, csv)
    for kv in a.option:
        k, v = kv.split("=")
        ddc.set_option(k, int(v))
    base = synth.digitiser_stream_fast(n, 1234, block=min(n, 1 << 22))
    if a.packed:
        row = torch.from_numpy(synth.pack10(base))
    else:
        row = torch.from_numpy(base.astype(np.float32))
    x = row.cuda().unsqueeze(0).repeat(a.streams, 1).contiguous()
    m = ddc.out_len(n)
    out = torch.empty((a.streams, m), dtype=torch.complex64, device="cuda")
    for _ in range(a.warmup):
        ddc.run_tensor(x, 100e6, out=out, packed=a.packed)
    torch.cuda.synchronize()
    if a.dbg:
        ddc.set_option("dbg_counters", 1)
    ts = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ddc.run_tensor(x, 100e6, out=out, packed=a.packed)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    if a.dbg:
        ddc.set_option("dbg_counters", 2)
    ms = float(np.median(ts))
    tot = a.streams * n
    bytes_ = tot * ((1.25 if a.packed else 4.0) + 8.0 / a.decim)
    flops = 4.0 * a.taps * m * a.streams
    print(f"T={a.taps} D={a.decim} N={n} streams={a.streams} packed={a.packed} variant={ddc.last_variant} "
          f"median={ms:.4f} ms min={min(ts):.4f} ms  {tot / ms / 1e6:.1f} Gsamples/s  {bytes_ / ms / 1e6:.0f} GB/s "
          f"({bytes_ / ms / 1e6 / 6539.5 * 100:.1f}% of 6539.5)  {flops / ms / 1e9:.2f} TFLOP/s ({flops / ms / 1e9 / 74.4 * 100:.1f}% of 74.4)")
    if a.check:
        from oracle import ddc_oracle as orc

        y = out[0].cpu().numpy()
        xin = base.astype(np.float32)
        step = orc.phase_step_cycles(n, 100e6, 1712e6)
        rng = np.random.default_rng(0)
        worst = 0.0
        for s0 in [0, m - 256] + [int(v) for v in rng.integers(0, m - 256, size=8)]:
            ref = orc.ddc_windowed_f64(xin, s0, 256, step, ddc.ddc_filter_coeffs, a.decim)
            worst = max(worst, float(np.abs(y[s0:s0 + 256] - ref).max() / np.abs(ref).max()))
        print(f"check: worst window max-err/max|ref| = {worst:.3e}")


if __name__ == "__main__":
    main()
