"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports /root/reference/feng/ddc/src/{ddc,cwg}.py, runs DigitalDownConverter.run on seeded inputs and
stores inputs + outputs as small .npz files; nothing here is needed at test time except the files it wrote.
"""
import io
import json
import os
import sys
import contextlib

import numpy as np

REF_SRC = "/root/reference/feng/ddc/src"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF_SRC)
import ddc as ref_ddc  # noqa: E402
import cwg as ref_cwg  # noqa: E402


def synth(n, seed, fs=1712e6, f0=103.3e6, amp=100.0, sigma=40.0):
    """Same generator as dc_sand_b200.synth.digitiser_stream (kept separate on purpose: fixtures must not
    change if the package changes)."""
    rng = np.random.default_rng(seed)
    phi = rng.uniform(0, 2 * np.pi)
    t = np.arange(n, dtype=np.float64)
    x = amp * np.cos(2 * np.pi * (f0 / fs) * t + phi) + sigma * rng.standard_normal(n)
    return np.clip(np.rint(x), -512, 511).astype(np.int16)


def make_ref(d, fs, csv):
    with contextlib.redirect_stdout(io.StringIO()):
        return ref_ddc.DigitalDownConverter(decimation_factor=d, sampling_frequency=fs, ddc_coeff_filename=csv)


def main():
    # ---- taps: integer numerators / 2^17, text form '%.5g' ---------------------------------------------
    taps_q17 = {}
    for name in ("ddc_coeff_107MHz.csv", "ddc_coeff_53MHz.csv"):
        txt = open(os.path.join(REF_SRC, name)).read()
        vals = np.array([float(s) for s in txt.split()])
        k = np.rint(vals * 2**17).astype(int)
        regen = "".join(" %.5g\n" % (kk / 2**17) if False else "%.5g\n" % (kk / 2**17) for kk in k)
        # byte-for-byte check modulo leading blanks / line ends
        assert [s for s in txt.split()] == regen.split(), name
        assert np.array_equal(np.array([float(s) for s in regen.split()]), vals)
        taps_q17[name] = [int(v) for v in k]
        taps_q17[name + ":raw_has_leading_space"] = txt.startswith(" ")
    json.dump(taps_q17, open(os.path.join(HERE, "taps_q17.json"), "w"))

    fs = 1712e6
    csv107 = os.path.join(REF_SRC, "ddc_coeff_107MHz.csv")
    csv53 = os.path.join(REF_SRC, "ddc_coeff_53MHz.csv")
    cases = [
        # name, N, csv, D, fc, seed
        ("n16384_d16", 16384, csv107, 16, 100e6, 1),
        ("n20001_d32_53", 20001, csv53, 32, 53.5e6, 2),
        ("n4099_d4", 4099, csv107, 4, 100e6, 3),
        ("n300_d16", 300, csv107, 16, 100e6, 4),
        ("n256_d16_single_output", 256, csv107, 16, 100e6, 5),
        ("n100_d16_swapped", 100, csv107, 16, 100e6, 6),
        ("n1_d16_swapped", 1, csv107, 16, 100e6, 7),
        ("n255_d16_swapped", 255, csv107, 16, 214e6, 8),
        ("n8192_d8_above_nyquist", 8192, csv107, 8, 1000e6, 9),
        ("n5000_d3_odd", 5000, csv107, 3, 100e6, 10),
        ("n2048_d1", 2048, csv53, 1, 428e6, 11),
        ("n70000_d64", 70000, csv53, 64, 100e6, 12),
        ("n40000_d16_fc214", 40000, csv107, 16, 214e6, 13),
    ]
    out = {}
    meta = {}
    for name, n, csv, d, fc, seed in cases:
        x = synth(n, seed)
        y = make_ref(d, fs, csv).run(x.astype(np.float32), fc)
        out[name + ":x"] = x
        out[name + ":y"] = np.asarray(y)
        meta[name] = dict(n=n, csv=os.path.basename(csv), d=d, fc=fc, fs=fs, seed=seed, m=int(len(y)))
        print(name, len(y), y.dtype)
    # a non-integer float32 input (tone produced by the reference's own generator, as its tests do)
    tone = ref_cwg.generate_carrier_wave(
        cw_scale=1, freq=103343750, sampling_frequency=fs, num_samples=32768, noise_scale=0, complex=False
    )
    assert tone.dtype == np.float32
    y = make_ref(16, fs, csv107).run(tone, 100e6)
    out["tone32768_d16:xf"] = tone
    out["tone32768_d16:y"] = np.asarray(y)
    meta["tone32768_d16"] = dict(n=32768, csv="ddc_coeff_107MHz.csv", d=16, fc=100e6, fs=fs, m=int(len(y)), float_input=True)
    np.savez_compressed(os.path.join(HERE, "ddc_small.npz"), **out)

    # ---- config 1 (BASELINE.json configs[0]): N = 2^20, T = 256, D = 16; store a strided subsample ---------
    n = 1 << 20
    x = synth(n, 1234)
    y = np.asarray(make_ref(16, fs, csv107).run(x.astype(np.float32), 100e6))
    meta["c1"] = dict(n=n, csv="ddc_coeff_107MHz.csv", d=16, fc=100e6, fs=fs, seed=1234, m=int(len(y)),
                      stride=64, sum_re=float(y.real.sum()), sum_im=float(y.imag.sum()),
                      l2=float(np.sqrt((np.abs(y) ** 2).sum())), max_abs=float(np.abs(y).max()))
    np.savez_compressed(os.path.join(HERE, "ddc_c1_subsample.npz"), y_sub=y[::64], y_head=y[:256], y_tail=y[-256:])

    # ---- the reference's own known-answer tests (feng/ddc/testing/test_ddc.py), recorded through the reference
    fft_length = 2**15
    n = fft_length * 16 * 2
    ka = {}

    def tone_(f):
        return ref_cwg.generate_carrier_wave(cw_scale=1, freq=f, sampling_frequency=fs, num_samples=n,
                                             noise_scale=0, complex=False)

    scen = {
        "center": ([100e6], 100e6, 1e5),
        "dual": ([100e6, 103343750], 100e6, 1e5),
        "bandedge": ([51019287.109375, 148980712.890625], 100e6, 1e5),
        "out_of_band": ([100e6, 214e6], 100e6, 1e3),
    }
    for k, (freqs, fc, thr) in scen.items():
        data = sum(tone_(f) for f in freqs)
        y = make_ref(16, fs, csv107).run(data, fc)
        p = np.abs(np.power(np.fft.fft(y[-fft_length:]), 2))
        bins = np.where(p > thr)[0]
        srt = np.sort(p)
        ka[k] = dict(freqs=freqs, fc=fc, threshold=thr, bins=[int(b) for b in bins], peak=float(srt[-1]),
                     second=float(srt[-2]), rejection_db=float(10 * np.log10(srt[-1] / srt[-2])))
        print(k, ka[k]["bins"], ka[k]["rejection_db"])
    # cwg known answers (feng/ddc/testing/test_cwg.py)
    cw = ref_cwg.generate_carrier_wave(cw_scale=1, freq=100e6, sampling_frequency=fs, num_samples=8192, noise_scale=0, complex=False)
    f = np.fft.rfft(np.real(cw))
    ka["cwg_real_bin"] = int(np.where(f == np.max(f))[0][0])
    cwc = ref_cwg.generate_carrier_wave(cw_scale=1, freq=214e6, sampling_frequency=fs, num_samples=8192, noise_scale=0, complex=True)
    f = np.fft.fft(cwc)
    ka["cwg_complex_bin"] = int(np.where(f == np.max(f))[0][0])
    out2 = {"cwg_real_8192": cw, "cwg_complex_8192": cwc}
    np.savez_compressed(os.path.join(HERE, "cwg_small.npz"), **out2)
    meta["known_answers"] = ka
    json.dump(meta, open(os.path.join(HERE, "meta.json"), "w"), indent=1)
    print("done")


if __name__ == "__main__":
    main()
