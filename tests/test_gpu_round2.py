"""Round-2 GPU parity tests ("next round" item 1): the configurations the first round left untested.

  * every FUSED packed-10-bit kernel fed the whole code range [-512, 511] (the first round's packed tests only used tone +
    noise, codes within +-260, so the sign-bit edge of the mantissa trick never reached the fused path);
  * BASELINE configs[2] at full size with oracle windows on many streams;
  * the single-process multi-device driver (threads, one handle per worker);
  * run() on the input dtypes / argument corners the reference accepts (vectors from the unmodified reference).
Tolerance as in conftest.py: max|y - y_ref| <= 1e-5 max|y_ref|, relative L2 <= 2e-6 (x4 above 256 taps); integer stages bit-exact."""
import os

import numpy as np
import pytest

from conftest import TOL_L2, TOL_MAX, rel_err

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from dc_sand_b200 import DigitalDownConverter, cwg as mycwg, synth, taps  # noqa: E402
from oracle import ddc_oracle as orc  # noqa: E402

FS = 1712e6


def _ddc(taps_dir, d, csv="ddc_coeff_107MHz.csv"):
    return DigitalDownConverter(decimation_factor=d, sampling_frequency=FS, ddc_coeff_filename=os.path.join(taps_dir, csv))


def _custom_taps(tmp_path, t):
    p = os.path.join(str(tmp_path), "taps.csv")
    np.savetxt(p, t, fmt="%.18e")
    return p


def _full_range_streams(n):
    """Two streams that exercise every 10-bit code at every position of the 4-sample / 5-byte group: (0) uniformly random
    codes over [-512, 511]; (1) all 1024 codes ascending at one group position and descending at another, the position
    advancing every 4096 samples, extremes (-512, 511, -1, 0) alternating in the other two."""
    rng = np.random.default_rng(2026)
    a = rng.integers(-512, 512, size=n).astype(np.int16)
    b = np.zeros(n, dtype=np.int16)
    codes = np.arange(-512, 512, dtype=np.int16)
    ext = np.array([-512, 511, -1, 0], dtype=np.int16)
    for blk in range(n // 4096):
        ph = blk % 4
        seg = b[blk * 4096:(blk + 1) * 4096]
        seg[ph::4] = np.roll(codes, 37 * blk)
        seg[(ph + 2) % 4::4] = np.roll(codes[::-1], 11 * blk)
        seg[(ph + 1) % 4::4] = ext[(np.arange(1024) + blk) % 4]
        seg[(ph + 3) % 4::4] = ext[(np.arange(1024) + blk + 2) % 4]
    return np.stack([a, b])


# CUDA-core engine for packed input (option packed_engine = 0): (D, T, variant option, expected kernel name fragment,
# bit-equal to the float32 kernel of the same loop?).  The in-warp-unpack kernels of round 1 (variants 5 and 7 on packed input)
# were superseded by the tensor engine and removed; what is left is the warp-specialised kernel at D = 16, 129 .. 256 taps and
# the bit-exact unpack stage in front of the float32 kernel of the cell everywhere else.
FUSED_PACKED = [
    (16, 256, 0, "packed10_split", True),      # warp-specialised: unpack warps -> float ring -> the float32 kernel's FIR loop
    (16, 256, 10, "packed10_split", True),
    (16, 200, 0, "packed10_split", False),     # 13 tap blocks padded to 16
    (16, 100, 0, "unpack10+", False),          # 7 tap blocks: unpack stage + row-staged float32 kernel
    (16, 1024, 0, "unpack10+", False),
    (8, 256, 0, "unpack10+", False),
    (32, 256, 0, "unpack10+", False),
    (64, 1000, 0, "unpack10+", False),
]


@pytest.mark.parametrize("d,t,variant,name,bit_equal", FUSED_PACKED)
def test_fused_unpack_kernels_full_code_range(d, t, variant, name, bit_equal, tmp_path):
    """north_star: bit-exact unpack, CUDA-core engine.  The fused kernel never materialises the unpacked samples, so exactness
    is shown through the outputs: with the FIR loop shared (warp-specialised kernel vs the float32 fast-FIR kernel) the packed
    and the float32 runs must be bit-identical; everywhere they must agree with the reference arithmetic on the unpacked samples."""
    from scipy import signal

    n = 64 * 4096 + 4096 + 320        # several chunks per CTA row, ragged last chunk, rows stay 16-byte aligned (n % 64 == 0)
    xs = _full_range_streams(n)
    packed = np.stack([orc.pack10(r) for r in xs])
    assert np.array_equal(orc.unpack10(packed[1]), xs[1])
    tp = signal.firwin(t, 0.8 / d) if t != 256 or d != 16 else taps.coefficients("ddc_coeff_107MHz.csv")
    ddc = DigitalDownConverter(d, FS, _custom_taps(tmp_path, tp))
    ddc.set_option("variant", variant)
    ddc.set_option("packed_engine", 0)   # the CUDA-core kernels (the tensor engine, default since round 2: test_gpu_tensor_engine.py)
    yp = ddc.run_tensor(torch.from_numpy(packed).cuda(), 100e6, packed=True)
    assert name in ddc.last_variant and "tensor_fir" not in ddc.last_variant, ddc.last_variant
    ddc.set_option("variant", 0)
    yf = ddc.run_tensor(torch.from_numpy(xs.astype(np.float32)).cuda(), 100e6)
    if bit_equal:
        assert "fast_fir<" in ddc.last_variant, ddc.last_variant
        assert torch.equal(yp, yf)
    k = 4 if t > 256 else 1
    yp = yp.cpu().numpy()
    scale = np.abs(yp).max()
    assert np.abs(yp - yf.cpu().numpy()).max() <= k * TOL_MAX * scale
    for s in range(2):
        ref = orc.ddc_reference(xs[s].astype(np.float32), 100e6, tp, d, FS)
        emax, el2 = rel_err(yp[s], ref)
        assert emax <= k * TOL_MAX and el2 <= k * TOL_L2, (s, ddc.last_variant, emax, el2)


def test_device_packer_is_inverse_of_unpack(taps_dir):
    """ddcb200_pack10 (test vectors built in HBM) against the oracle's packer, all codes, incl. rounding and clipping."""
    xs = _full_range_streams(8 * 4096).astype(np.float32)
    got = mycwg.pack10_gpu(torch.from_numpy(xs).cuda()).cpu().numpy()
    assert np.array_equal(got, np.stack([orc.pack10(r.astype(np.int16)) for r in xs]))
    odd = np.array([511.4, 511.6, 600.0, -512.4, -513.0, -0.5, 0.5, 1.5], dtype=np.float32)   # rint (ties to even) then clip
    want = np.clip(np.rint(odd), -512, 511).astype(np.int16)
    assert np.array_equal(orc.unpack10(mycwg.pack10_gpu(torch.from_numpy(odd).cuda()).cpu().numpy()), want)
    ddc = _ddc(taps_dir, 16)
    assert np.array_equal(ddc._decode_8bit_to_10bit_to_float_data(got[1]), xs[1])


def test_packed_config_full_size_many_streams(taps_dir):
    """BASELINE configs[2] at full size (64 streams x 2^24 packed samples): every stream has its own content (generated in
    HBM: digitiser model, per-stream seed), packed on the device, and windows of the result -- head, tail and random, on ten
    streams including the first and the last -- are checked against the float64 windowed oracle fed with the UNPACKED bytes
    of exactly the window that was read back."""
    n, s, d, t = 1 << 24, 64, 16, 256
    ddc = _ddc(taps_dir, d)
    xf = mycwg.generate_carrier_wave_gpu(100.0, 103.3e6, FS, n, 40.0, False, seed=77, n_streams=s, digitise=True)
    xp = mycwg.pack10_gpu(xf)
    del xf
    yp = ddc.run_tensor(xp, 100e6, packed=True)
    assert "packed10" in ddc.last_variant and "unpack10+" not in ddc.last_variant and yp.shape == (s, ddc.out_len(n))
    m = yp.shape[1]
    scale = float(yp.abs().max())
    step = orc.phase_step_cycles(n, 100e6, FS)
    rng = np.random.default_rng(3)
    worst = 0.0
    for k in (0, 1, 7, 13, 22, 31, 32, 47, 62, 63):
        for m0 in [0, m - 512] + [int(v) // 4 * 4 for v in rng.integers(0, m - 512, size=2)]:   # m0 D / 4 * 5 stays a whole byte
            b0, b1 = m0 * d // 4 * 5, ((m0 + 511) * d + t) // 4 * 5
            seg = orc.unpack10(xp[k, b0:b1].cpu().numpy()).astype(np.float32)
            ref = orc.ddc_windowed_f64(seg, m0, 512, step, ddc.ddc_filter_coeffs, d, x_base=m0 * d)
            err = float(np.abs(yp[k, m0:m0 + 512].cpu().numpy() - ref).max())
            worst = max(worst, err)
            assert err <= TOL_MAX * scale, (k, m0, err / scale)
    print("C3 full size: worst window error / max|y| =", worst / scale)


def test_multi_device_driver_threads(taps_dir):
    """MultiDeviceDDC (one handle per worker, one Python thread per worker, GIL released in the C ABI).  With one GPU the two
    workers share device 0 -- still two handles driven concurrently from two threads, on configurations that need the
    tensor-map entry point (D = 8: tensor-staged kernel) the first time either thread runs; with several GPUs the streams
    are sharded over all of them."""
    from dc_sand_b200.scheduler import MultiDeviceDDC

    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev >= 2 else [0, 0]
    for d, csv in ((8, "ddc_coeff_107MHz.csv"), (16, "ddc_coeff_107MHz.csv"), (32, "ddc_coeff_53MHz.csv")):
        n, s = 300_000, 7
        x = np.stack([synth.digitiser_stream(n, 4000 + i) for i in range(s)]).astype(np.float32)
        md = MultiDeviceDDC(d, FS, os.path.join(taps_dir, csv), devices=devices)
        y = md.run_batch(x, 100e6)
        assert y.shape == (s, md.workers[0].out_len(n)) and y.dtype == np.complex64
        for i in range(s):
            ref = orc.ddc_reference(x[i], 100e6, md.workers[0].ddc_filter_coeffs, d, FS)
            emax, el2 = rel_err(y[i], ref)
            assert emax <= TOL_MAX and el2 <= TOL_L2, (d, i, emax, el2)
        md.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_multi_device_driver_two_gpus(taps_dir):
    """Streams sharded over two real devices: every shard equals a single-device run of the same streams."""
    from dc_sand_b200.scheduler import MultiDeviceDDC

    n, s = 1 << 20, 6
    x = np.stack([synth.digitiser_stream_fast(n, 4100 + i) for i in range(s)]).astype(np.float32)
    md = MultiDeviceDDC(16, FS, os.path.join(taps_dir, "ddc_coeff_107MHz.csv"), devices=[0, 1])
    y = md.run_batch(x, 100e6)
    one = _ddc(taps_dir, 16)
    y1 = one.run_batch(x, 100e6)
    assert np.array_equal(y, y1)
    md.close()


def test_run_accepts_the_reference_input_types(meta_r2, golden_r2, taps_dir):
    """int16 / int8 / float64 / complex input, negative centre frequency, full-range codes: run() against outputs of the
    unmodified reference for exactly these inputs (tests/golden/make_golden_r2.py)."""
    for name, m in meta_r2.items():
        x = golden_r2[name + ":x"]
        y_ref = golden_r2[name + ":y"]
        ddc = _ddc(taps_dir, m["d"], m["csv"])
        y = ddc.run(x, m["fc"])
        assert y.dtype == np.complex128 and y.shape == y_ref.shape, name
        emax, el2 = rel_err(y, y_ref)
        assert emax <= TOL_MAX and el2 <= TOL_L2, (name, emax, el2, ddc.last_variant)


def test_stream_sessions_are_tied_to_their_converter(taps_dir):
    """A DDCStream holds the native handle of its converter: closing the converter closes its sessions first, and taps /
    decimation cannot change under an open session (its carry and buffers were sized for the old geometry)."""
    from dc_sand_b200 import DDCStream

    x = synth.digitiser_stream(50_000, 1).astype(np.float32)
    ddc = _ddc(taps_dir, 16)
    st = DDCStream(ddc, 100e6)
    st.push(x[:20_000])
    ddc.decimation_factor = 8
    with pytest.raises(RuntimeError, match="DDCStream"):
        ddc.run(x, 100e6)
    ddc.decimation_factor = 16
    ddc.close()                      # closes the session, then the handle
    assert st._s is None
    st.close()                       # idempotent
    ddc2 = _ddc(taps_dir, 16)
    ddc2.decimation_factor = 2.5
    with pytest.raises(ValueError, match="positive integer"):
        ddc2.run(x, 100e6)
