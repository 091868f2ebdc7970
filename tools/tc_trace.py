#!/usr/bin/env python
"""Per-tile event trace of the tensor engine (CTA 0, first 96 tiles) on BASELINE configs[2]: needs a library built with
-DDDCB200_TC_TRACE (e.g. `make -C dc_sand_b200/csrc EXTRA=-DDDCB200_TC_TRACE OUT=/tmp/libtrace.so OBJDIR=/tmp/trace_obj`, then
DDCB200_LIB=/tmp/libtrace.so python tools/tc_trace.py 2> trace.txt).  Columns: clocks since the first event."""
import os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, ".")
from dc_sand_b200 import DigitalDownConverter, synth, taps
ddc = DigitalDownConverter(16, 1712e6, taps.write_csv("ddc_coeff_107MHz.csv", tempfile.mkdtemp()))
n, S = 1 << 24, 64
base = synth.digitiser_stream_fast(n, 1, block=1 << 22)
x = torch.from_numpy(synth.pack10(base)).cuda().unsqueeze(0).repeat(S, 1).contiguous()
out = torch.empty((S, ddc.out_len(n)), dtype=torch.complex64, device="cuda")
for _ in range(3): ddc.run_tensor(x, 100e6, out=out, packed=True)
torch.cuda.synchronize()
ddc.set_option("dbg_counters", 1)
ddc.run_tensor(x, 100e6, out=out, packed=True)
torch.cuda.synchronize()
ddc.set_option("dbg_counters", 3)
