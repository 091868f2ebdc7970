#!/usr/bin/env python
"""Back-to-back kernel timing (no host sync between launches):
    python tools/b2b.py [--iters=200] [packed=1] [streams=S] [logn=L] [taps=T] [decim=D] [dbg=1] key=value ...
key=value pairs other than those go to ddcb200_set_option (variant, debug_mode, ...)."""
import os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dc_sand_b200 import DigitalDownConverter, synth, taps
iters = 200
opts, cfg = [], {"packed": 0, "streams": 1, "logn": 28, "taps": 256, "decim": 16, "dbg": 0}
for a in sys.argv[1:]:
    if a.startswith("--iters="): iters = int(a.split("=")[1])
    elif a.split("=")[0] in cfg: cfg[a.split("=")[0]] = int(a.split("=")[1])
    else: opts.append(a)
tmp = tempfile.mkdtemp()
T, D = cfg["taps"], cfg["decim"]
if T == 256:
    csv = taps.write_csv("ddc_coeff_107MHz.csv", tmp)
else:
    from scipy import signal
    csv = os.path.join(tmp, "t.csv"); np.savetxt(csv, signal.firwin(T, 0.8 / D), fmt="%.18e")
ddc = DigitalDownConverter(D, 1712e6, csv)
for kv in opts:
    k, v = kv.split("="); ddc.set_option(k, int(v))
n, S, packed = 1 << cfg["logn"], cfg["streams"], bool(cfg["packed"])
base = synth.digitiser_stream_fast(n, 1, block=min(n, 1 << 22))
row = torch.from_numpy(synth.pack10(base) if packed else base.astype(np.float32))
x = row.cuda().unsqueeze(0).repeat(S, 1).contiguous()
out = torch.empty((S, ddc.out_len(n)), dtype=torch.complex64, device="cuda")
for _ in range(5): ddc.run_tensor(x, 100e6, out=out, packed=packed)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ddc.run_tensor(x, 100e6, out=out, packed=packed)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / iters)
tot = n * S
b = tot * ((1.25 if packed else 4.0) + 8.0 / D)
print(f"{' '.join(sys.argv[1:]):44s} {ddc.last_variant}: {best:.4f} ms  {tot / best / 1e6:.1f} Gsamples/s  {b / best / 1e6:.0f} GB/s "
      f"({b / best / 1e6 / 65.395:.1f}% HBM)  {4 * T * ddc.out_len(n) * S / best / 1e9 / 74.4 * 100:.1f}% FP32(direct-form flops)")
if cfg["dbg"]:
    ddc.set_option("dbg_counters", 1)
    for _ in range(20): ddc.run_tensor(x, 100e6, out=out, packed=packed)
    torch.cuda.synchronize()
    ddc.set_option("dbg_counters", 2)
