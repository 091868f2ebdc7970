"""Round-2 golden fixtures from the UNMODIFIED reference: the input dtypes and argument corners `run()` accepts beyond
float32 (int16, int8, float64 with non-representable values, complex input, negative centre frequency).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_r2.py

Writes tests/golden/ddc_r2.npz + meta_r2.json; nothing here is needed at test time except those files.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np

REF_SRC = "/root/reference/feng/ddc/src"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF_SRC)
import ddc as ref_ddc  # noqa: E402


def synth(n, seed, fs=1712e6, f0=103.3e6, amp=100.0, sigma=40.0):
    rng = np.random.default_rng(seed)
    phi = rng.uniform(0, 2 * np.pi)
    t = np.arange(n, dtype=np.float64)
    x = amp * np.cos(2 * np.pi * (f0 / fs) * t + phi) + sigma * rng.standard_normal(n)
    return np.clip(np.rint(x), -512, 511).astype(np.int16)


def make_ref(d, fs, csv):
    with contextlib.redirect_stdout(io.StringIO()):
        return ref_ddc.DigitalDownConverter(decimation_factor=d, sampling_frequency=fs, ddc_coeff_filename=csv)


def main():
    fs = 1712e6
    csv107 = os.path.join(REF_SRC, "ddc_coeff_107MHz.csv")
    csv53 = os.path.join(REF_SRC, "ddc_coeff_53MHz.csv")
    rng = np.random.default_rng(99)
    out, meta = {}, {}

    def case(name, x, csv, d, fc):
        with contextlib.redirect_stdout(io.StringIO()):
            y = make_ref(d, fs, csv).run(x, fc)
        out[name + ":x"] = x
        out[name + ":y"] = y
        meta[name] = {"n": int(len(x)), "csv": os.path.basename(csv), "d": d, "fc": fc, "fs": fs, "m": int(len(y)),
                      "in_dtype": str(x.dtype), "out_dtype": str(y.dtype)}

    x = synth(30000, 21)
    case("int16_in", x, csv107, 16, 100e6)                                  # integer array handed over as is
    case("int8_in", np.clip(x // 4, -128, 127).astype(np.int8), csv107, 16, 100e6)
    xf = synth(24000, 22).astype(np.float64) + rng.uniform(-0.5, 0.5, 24000)   # float64 values float32 cannot hold
    case("float64_in", xf, csv107, 8, 214e6)
    case("neg_fc", synth(20000, 23).astype(np.float32), csv107, 16, -100e6)
    case("neg_fc_53", synth(33000, 24).astype(np.float32), csv53, 32, -53.5e6)
    xc = (synth(16000, 25).astype(np.float32) + 1j * synth(16000, 26).astype(np.float32)).astype(np.complex64)
    case("complex_in", xc, csv107, 16, 100e6)
    case("uniform_full_range", rng.integers(-512, 512, size=40000).astype(np.float32), csv107, 16, 100e6)

    np.savez_compressed(os.path.join(HERE, "ddc_r2.npz"), **out)
    json.dump(meta, open(os.path.join(HERE, "meta_r2.json"), "w"), indent=1)
    for k, v in meta.items():
        print(k, v)


if __name__ == "__main__":
    main()
