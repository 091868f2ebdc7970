import os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dc_sand_b200 import DigitalDownConverter, synth, taps
ddc = DigitalDownConverter(16, 1712e6, taps.write_csv("ddc_coeff_107MHz.csv", tempfile.mkdtemp()))
for kv in sys.argv[1:]:
    k, v = kv.split("="); ddc.set_option(k, int(v))
n = 1 << 28
x = torch.from_numpy(synth.digitiser_stream_fast(n, 1, block=1 << 22).astype(np.float32)).cuda().unsqueeze(0)
out = torch.empty((1, ddc.out_len(n)), dtype=torch.complex64, device="cuda")
for _ in range(5): ddc.run_tensor(x, 100e6, out=out)
torch.cuda.synchronize()
ddc.set_option("dbg_counters", 1)
for _ in range(20): ddc.run_tensor(x, 100e6, out=out)
torch.cuda.synchronize()
ddc.set_option("dbg_counters", 2)
