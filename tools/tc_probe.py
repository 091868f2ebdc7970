#!/usr/bin/env python
"""Parity probe of the tensor-core packed engine (option variant=13) against the float64 oracle and the CUDA-core kernels.

    python tools/tc_probe.py [--cases small|sweep|full]
"""
import argparse
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dc_sand_b200 import DigitalDownConverter, synth, taps  # noqa: E402
from oracle import ddc_oracle as orc  # noqa: E402


OPTIONS = []   # (key, value) pairs from --option, passed to ddcb200_set_option (e.g. tc_mode=1)


def make_ddc(T, D, tmp):
    if T == 256 and D == 16:
        csv = taps.write_csv("ddc_coeff_107MHz.csv", tmp)
    else:
        from scipy import signal

        csv = os.path.join(tmp, f"t{T}_{D}.csv")
        np.savetxt(csv, signal.firwin(T, 0.8 / D), fmt="%.18e")
    return DigitalDownConverter(D, 1712e6, csv)


def one(T, D, n, streams, tmp, fc=100e6, full_range=False, verbose=True):
    ddc = make_ddc(T, D, tmp)
    rng = np.random.default_rng(T * 131 + D)
    rows = []
    for s in range(streams):
        if full_range:
            rows.append(rng.integers(-512, 512, size=n).astype(np.int16))
        else:
            rows.append(synth.digitiser_stream_fast(n, 100 + s, block=min(n, 1 << 20)))
    packed = np.stack([synth.pack10(r) for r in rows])
    x = torch.from_numpy(packed).cuda()
    m = ddc.out_len(n)
    out = torch.zeros((streams, m), dtype=torch.complex64, device="cuda")
    ddc.set_option("variant", 13)
    for k, v in OPTIONS:
        ddc.set_option(k, v)
    ddc.run_tensor(x, fc, out=out, packed=True)
    torch.cuda.synchronize()
    name = ddc.last_variant
    y = out.cpu().numpy()
    step = orc.phase_step_cycles(n, fc, 1712e6)
    worst = 0.0
    worst_l2 = 0.0
    for s in range(streams):
        xin = rows[s].astype(np.float64)
        if m <= 8192:
            ref = orc.ddc_windowed_f64(xin, 0, m, step, ddc.ddc_filter_coeffs, D)
            e = np.abs(y[s] - ref)
            worst = max(worst, float(e.max() / np.abs(ref).max()))
            worst_l2 = max(worst_l2, float(np.linalg.norm(y[s] - ref) / np.linalg.norm(ref)))
        else:
            r2 = np.random.default_rng(s)
            for s0 in [0, m - 512] + [int(v) for v in r2.integers(0, m - 512, size=4)]:
                ref = orc.ddc_windowed_f64(xin, s0, 512, step, ddc.ddc_filter_coeffs, D)
                e = np.abs(y[s, s0:s0 + 512] - ref)
                worst = max(worst, float(e.max() / np.abs(ref).max()))
                worst_l2 = max(worst_l2, float(np.linalg.norm(y[s, s0:s0 + 512] - ref) / np.linalg.norm(ref)))
    if verbose:
        print(f"T={T} D={D} n={n} streams={streams} full_range={full_range} {name}: max_err={worst:.3e} rel_l2={worst_l2:.3e}", flush=True)
    ddc.close()
    return worst, worst_l2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="small")
    ap.add_argument("--option", action="append", default=[], help="key=value passed to ddcb200_set_option")
    a = ap.parse_args()
    OPTIONS.extend((kv.split("=")[0], int(kv.split("=")[1])) for kv in a.option)
    tmp = tempfile.mkdtemp()
    if a.cases == "small":
        one(256, 16, 8192 + 256, 1, tmp)
        one(256, 16, 40000, 3, tmp)
        one(256, 16, 1 << 20, 2, tmp, full_range=True)
    elif a.cases == "sweep":
        for D in (4, 8, 16, 32, 64):
            for T in (64, 128, 256, 512, 1024):
                if T < D:
                    continue
                try:
                    one(T, D, 300032, 2, tmp)
                except Exception as e:  # unsupported cells raise
                    print(f"T={T} D={D}: {e}")
    elif a.cases == "full":
        one(256, 16, 1 << 24, 4, tmp)


if __name__ == "__main__":
    main()
