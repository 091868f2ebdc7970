#!/usr/bin/env python
"""Runs the fused DDC back to back for a few seconds while sampling SM clock / power / throttle reasons (NVML)."""
import os, sys, tempfile, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dc_sand_b200 import DigitalDownConverter, synth, taps
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
ddc = DigitalDownConverter(16, 1712e6, taps.write_csv("ddc_coeff_107MHz.csv", tempfile.mkdtemp()))
for kv in sys.argv[1:]:
    k, v = kv.split("="); ddc.set_option(k, int(v))
n = 1 << 28
x = torch.from_numpy(synth.digitiser_stream_fast(n, 1, block=1 << 22).astype(np.float32)).cuda().unsqueeze(0)
out = torch.empty((1, ddc.out_len(n)), dtype=torch.complex64, device="cuda")
samples = []; stop = threading.Event()
def loop():
    while not stop.is_set():
        samples.append((time.time(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.01)
t = threading.Thread(target=loop); t.start()
time.sleep(0.2)
for rep in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record()
    for _ in range(1500): ddc.run_tensor(x, 100e6, out=out)
    e1.record(); torch.cuda.synchronize(); t1 = time.time()
    busy = [s for s in samples if t0 + 0.05 < s[0] < t1]
    print(f"rep {rep}: {e0.elapsed_time(e1) / 1500:.4f} ms/iter  sm_clk median {np.median([s[1] for s in busy]):.0f} min {min(s[1] for s in busy)} max {max(s[1] for s in busy)} MHz  "
          f"power median {np.median([s[2] for s in busy]):.0f} W  reasons {sorted(set(hex(s[3]) for s in busy))}  variant {ddc.last_variant}")
stop.set(); t.join()
