"""GPU parity tests of the tensor-core engine for packed 10-bit input (dc_sand_b200/csrc/ddc_kernel_tc.cuh), the default
engine for packed input since round 2.

What has to hold (north_star): the 10-bit unpack is bit-exact and the complex baseband output matches the reference's NumPy
DDC (feng/ddc/src/ddc.py:121-162 on the unpacked samples) within the repo-wide tolerance of conftest.py: max|y - y_ref| <=
1e-5 max|y_ref| and relative L2 <= 2e-6 (x4 above 256 taps).  The engine never materialises the unpacked samples, so the
exactness of the unpack is shown through outputs: every 10-bit code at every position of the 5-byte group, and an impulse
response test in which each output must equal ONE tap times ONE code (any wrong bit of the unpack moves it by >= 1 LSB of
the code, 2^-9 of full scale, four orders of magnitude above the tolerance)."""
import os

import numpy as np
import pytest

from conftest import TOL_L2, TOL_MAX, rel_err

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from dc_sand_b200 import DDCStream, DigitalDownConverter, synth, taps  # noqa: E402
from oracle import ddc_oracle as orc  # noqa: E402
from test_gpu_round2 import _custom_taps, _full_range_streams  # noqa: E402

FS = 1712e6


def _taps_for(t, d):
    from scipy import signal

    return taps.coefficients("ddc_coeff_107MHz.csv") if (t, d) == (256, 16) else signal.firwin(t, 0.8 / d)


def _run_packed(ddc, xs, fc=100e6, **kw):
    packed = np.stack([orc.pack10(r) for r in xs])
    return ddc.run_tensor(torch.from_numpy(packed).cuda(), fc, packed=True, **kw)


# (D, T): every decimation the engine is built for; 64- and 128-sample rows; R = 16, 8, 4 outputs per row (fragment-layout
# epilogue, one and two 32-column blocks, the 16-column block) and R = 2 (row-per-lane epilogue); short and long filters
CELLS = [(4, 64), (4, 300), (8, 128), (8, 1024), (16, 256), (16, 40), (16, 1024), (32, 256), (32, 700), (64, 512), (64, 64)]


@pytest.mark.parametrize("d,t", CELLS)
def test_tensor_engine_full_code_range(d, t, tmp_path):
    """Uniformly random codes over [-512, 511] and an all-1024-codes x 4-positions pattern (test_gpu_round2._full_range_streams)
    through the engine: against the reference arithmetic on the unpacked samples and against the float32 CUDA-core kernel."""
    n = 3 * 16384 + 4096 + 320        # a few tiles per stream and a ragged last one; rows stay 16-byte aligned (n % 64 == 0)
    xs = _full_range_streams(n)
    tp = _taps_for(t, d)
    ddc = DigitalDownConverter(d, FS, _custom_taps(tmp_path, tp))
    yp = _run_packed(ddc, xs)
    assert "tensor_fir_packed10" in ddc.last_variant, ddc.last_variant
    yf = ddc.run_tensor(torch.from_numpy(xs.astype(np.float32)).cuda(), 100e6)
    assert "tensor" not in ddc.last_variant or "staged" in ddc.last_variant, ddc.last_variant   # float32 stays on the CUDA cores
    k = 4 if t > 256 else 1
    yp = yp.cpu().numpy()
    scale = np.abs(yp).max()
    assert np.abs(yp - yf.cpu().numpy()).max() <= k * TOL_MAX * scale
    for s in range(2):
        ref = orc.ddc_reference(xs[s].astype(np.float32), 100e6, tp, d, FS)
        emax, el2 = rel_err(yp[s], ref)
        assert yp[s].shape == ref.shape and emax <= k * TOL_MAX and el2 <= k * TOL_L2, (s, ddc.last_variant, emax, el2)


@pytest.mark.parametrize("d", [8, 16, 64])
def test_tensor_engine_unpack_is_bit_exact(d, tmp_path):
    """Impulse-response form of "bit-exact unpack": one tap equal to 1, all others 0, and a centre frequency that makes the
    reference's NCO advance exactly one whole cycle per sample (cwg.py:31-33: int(N / (fs / f)) = N - 1 cycles over N - 1
    steps; f = 0 raises, as in the reference).  Output m is then exactly code x[D m + T - 1 - k0] (the NCO factor is 1, the tap
    splits hi + lo exactly, one product, nothing to round), so the whole output must equal the unpacked codes as integers --
    for every code value at every position of the byte group."""
    t = 64
    n = 5 * 16384 + 64
    xs = _full_range_streams(n)
    for k0 in (0, 17, t - 1):
        tp = np.zeros(t)
        tp[k0] = 1.0
        ddc = DigitalDownConverter(d, FS, _custom_taps(tmp_path, tp))
        fc = FS * (n - 0.5) / n
        assert ddc.phase_step(n, fc) == 1.0
        y = _run_packed(ddc, xs, fc=fc).cpu().numpy()
        assert "tensor_fir_packed10" in ddc.last_variant, ddc.last_variant
        m = y.shape[1]
        want = xs[:, t - 1 - k0::d][:, :m].astype(np.float64)     # convolution: out[m] = sum_k h[k] x[D m + T - 1 - k]
        assert np.array_equal(y.real.astype(np.float64), want), (d, k0, np.abs(y.real - want).max())
        assert not y.imag.any()
        ddc.close()


@pytest.mark.parametrize("n", [256, 320, 4096 + 64, 16384 - 64, 16384 + 192, 16384 + 256, 2 * 16384 + 64, 100_032, 4108, 100_004])
def test_tensor_engine_ragged_lengths(n, taps_dir):
    """Stream lengths around the tile size (16384 samples = 1024 outputs at 128-sample rows): a single output, shorter than
    one tile, a partly filled last row, one output into the next tile; rows whose byte length is not a multiple of 16
    (n % 64 != 0) make the row pitch unaligned, which the dispatcher must route to the CUDA-core kernels instead."""
    d, t = 16, 256
    ddc = DigitalDownConverter(d, FS, os.path.join(taps_dir, "ddc_coeff_107MHz.csv"))
    xs = np.stack([synth.digitiser_stream(n, 300 + s) for s in range(3)])
    y = _run_packed(ddc, xs).cpu().numpy()
    aligned = (n // 4 * 5) % 16 == 0
    assert ("tensor_fir_packed10" in ddc.last_variant) == aligned, (n, ddc.last_variant)
    ref = np.stack([orc.ddc_reference(r.astype(np.float32), 100e6, ddc.ddc_filter_coeffs, d, FS) for r in xs])
    emax, el2 = rel_err(y, ref)
    assert y.shape == ref.shape and emax <= TOL_MAX and el2 <= TOL_L2, (n, ddc.last_variant, emax, el2)


def test_tensor_engine_padded_rows_and_offset_outputs(taps_dir):
    """Rows inside a wider buffer (pitch a multiple of 16 bytes, longer than the row) and an output matrix with a pitch: the
    engine reads only each row's own bytes (the ragged tail of a row must not pick up its neighbour's pad) and writes only
    each row's own outputs."""
    d, n, s = 16, 50_000 - 50_000 % 64, 5
    ddc = DigitalDownConverter(d, FS, os.path.join(taps_dir, "ddc_coeff_107MHz.csv"))
    xs = np.stack([synth.digitiser_stream(n, 500 + k) for k in range(s)])
    row = n // 4 * 5
    pitch = row + 48
    buf = torch.full((s, pitch), 0xA5, dtype=torch.uint8, device="cuda")           # pad bytes that decode to non-zero codes
    buf[:, :row] = torch.from_numpy(np.stack([orc.pack10(r) for r in xs])).cuda()
    m = ddc.out_len(n)
    out = torch.full((s, m + 6), 7.0 + 7.0j, dtype=torch.complex64, device="cuda")
    ddc.run_tensor(buf[:, :row], 100e6, out=out[:, :m], packed=True)
    assert "tensor_fir_packed10" in ddc.last_variant, ddc.last_variant
    y = out.cpu().numpy()
    assert np.all(y[:, m:] == 7.0 + 7.0j)
    ref = np.stack([orc.ddc_reference(r.astype(np.float32), 100e6, ddc.ddc_filter_coeffs, d, FS) for r in xs])
    emax, el2 = rel_err(y[:, :m], ref)
    assert emax <= TOL_MAX and el2 <= TOL_L2, (emax, el2)


def test_tensor_engine_tap_reload_and_frequency_change(taps_dir, tmp_path):
    """The engine caches its tap-matrix image per (NCO step, decimation, taps): a new centre frequency, new taps at run time
    and a new decimation must each rebuild it; going back to an earlier key must give the earlier result bit for bit."""
    n = 3 * 16384
    xs = _full_range_streams(n)[:1]
    ddc = DigitalDownConverter(16, FS, os.path.join(taps_dir, "ddc_coeff_107MHz.csv"))
    first = _run_packed(ddc, xs, fc=100e6).cpu().numpy()
    for fc in (53.5e6, 428e6, 100e6):
        y = _run_packed(ddc, xs, fc=fc).cpu().numpy()
        ref = orc.ddc_reference(xs[0].astype(np.float32), fc, ddc.ddc_filter_coeffs, 16, FS)
        emax, el2 = rel_err(y[0], ref)
        assert emax <= TOL_MAX and el2 <= TOL_L2, (fc, emax, el2)
    assert np.array_equal(y, first)
    ddc.ddc_filter_coeffs = taps.coefficients("ddc_coeff_53MHz.csv")        # narrow-band mode, same length
    y = _run_packed(ddc, xs, fc=100e6).cpu().numpy()
    ref = orc.ddc_reference(xs[0].astype(np.float32), 100e6, ddc.ddc_filter_coeffs, 16, FS)
    emax, el2 = rel_err(y[0], ref)
    assert "tensor_fir_packed10" in ddc.last_variant and emax <= TOL_MAX and el2 <= TOL_L2, (ddc.last_variant, emax, el2)
    ddc.decimation_factor = 32
    y = _run_packed(ddc, xs, fc=100e6).cpu().numpy()
    ref = orc.ddc_reference(xs[0].astype(np.float32), 100e6, ddc.ddc_filter_coeffs, 32, FS)
    emax, el2 = rel_err(y[0], ref)
    assert "D32" in ddc.last_variant and emax <= TOL_MAX and el2 <= TOL_L2, (ddc.last_variant, emax, el2)


def test_tensor_engine_streaming_session(taps_dir):
    """Ragged packed pushes through a session (carry of T - D samples in front of every chunk, NCO phase continued through
    sample_offset) equal one run over the whole stream; device pushes land on 16-byte aligned rows, so they take the engine."""
    n = 400_000
    xi = np.stack([synth.digitiser_stream(n, 700 + s) for s in range(3)])
    packed = np.stack([synth.pack10(r) for r in xi])
    ddc = DigitalDownConverter(16, FS, os.path.join(taps_dir, "ddc_coeff_107MHz.csv"))
    ref = np.stack([orc.ddc_reference(r.astype(np.float32), 100e6, ddc.ddc_filter_coeffs, 16, FS) for r in xi])
    cuts = [0, 64, 65_600, 65_664, 200_000, 330_048, n]
    seen = set()
    with DDCStream(ddc, 100e6, n_streams=3, max_chunk=140_000, total_samples=n, packed=True) as st:
        parts = []
        for i, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
            piece = np.ascontiguousarray(packed[:, a // 4 * 5: b // 4 * 5])
            parts.append(st.push_tensor(torch.from_numpy(piece).cuda()).cpu().numpy() if i % 2 == 0 else st.push(piece))
            seen.add(ddc.last_variant.split("<")[0])
    y = np.concatenate(parts, axis=1)
    emax, el2 = rel_err(y, ref)
    assert y.shape == ref.shape and emax <= TOL_MAX and el2 <= TOL_L2, (y.shape, emax, el2, seen)
    print("session kernels:", sorted(seen))


def test_tensor_engine_selection(taps_dir, tmp_path):
    """Engine selection: default -> tensor engine; option packed_engine = 0 -> CUDA-core fused kernel; a decimation the engine
    is not built for (not a power of two) -> CUDA cores whatever the option; forced (variant 13) on float32 input -> ignored."""
    from scipy import signal

    n = 2 * 16384
    xs = _full_range_streams(n)[:1]
    ddc = DigitalDownConverter(16, FS, os.path.join(taps_dir, "ddc_coeff_107MHz.csv"))
    a = _run_packed(ddc, xs).cpu().numpy()
    assert "tensor_fir_packed10" in ddc.last_variant
    ddc.set_option("packed_engine", 0)
    b = _run_packed(ddc, xs).cpu().numpy()
    assert "tensor" not in ddc.last_variant and "packed10" in ddc.last_variant, ddc.last_variant
    assert np.abs(a - b).max() <= TOL_MAX * np.abs(b).max()
    ddc.set_option("packed_engine", 1)
    ddc.set_option("variant", 13)
    ddc.run_tensor(torch.from_numpy(xs.astype(np.float32)).cuda(), 100e6)
    assert "tensor_fir" not in ddc.last_variant, ddc.last_variant
    odd = DigitalDownConverter(12, FS, _custom_taps(tmp_path, signal.firwin(96, 0.8 / 12)))
    y = _run_packed(odd, xs).cpu().numpy()
    assert "tensor_fir" not in odd.last_variant, odd.last_variant
    ref = orc.ddc_reference(xs[0].astype(np.float32), 100e6, odd.ddc_filter_coeffs, 12, FS)
    emax, el2 = rel_err(y[0], ref)
    assert emax <= TOL_MAX and el2 <= TOL_L2, (emax, el2)


def test_tensor_engine_two_streams_of_work_share_a_handle(taps_dir):
    """Asynchronous calls on two torch streams with different centre frequencies through one handle: each call's tap-matrix
    image must stay alive until its kernel has run (ring of images guarded by events, k_tc.cu)."""
    n = 1 << 20
    xs = np.stack([synth.digitiser_stream_fast(n, 900 + s) for s in range(4)])
    xp = torch.from_numpy(np.stack([synth.pack10(r) for r in xs])).cuda()
    ddc = DigitalDownConverter(16, FS, os.path.join(taps_dir, "ddc_coeff_107MHz.csv"))
    fcs = [100e6, 53.5e6, 428e6, 1.0e6, 855e6, 100e6]
    want = [ddc.run_tensor(xp, fc, packed=True).clone() for fc in fcs]
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    got = [None] * len(fcs)
    for rep in range(3):
        for i, fc in enumerate(fcs):
            with torch.cuda.stream(s1 if i % 2 else s2):
                got[i] = ddc.run_tensor(xp, fc, packed=True)
        torch.cuda.synchronize()
        for i in range(len(fcs)):
            assert torch.equal(got[i], want[i]), (rep, i)


@pytest.mark.parametrize("packed", [False, True])
@pytest.mark.parametrize("n,s", [(65_536, 7), (50_003 * 4, 5), (1 << 20, 2)])
def test_host_path_by_stream_chunks_equal_time_chunks(n, s, packed, taps_dir):
    """The host entry points cut a batch of short streams BY STREAM (whole rows per chunk, flat copies) and long streams in
    TIME; both chunkings, several chunk budgets, contiguous and odd row lengths must give the one-shot device result."""
    ddc = DigitalDownConverter(16, FS, os.path.join(taps_dir, "ddc_coeff_107MHz.csv"))
    xs = np.stack([synth.digitiser_stream_fast(n, 40 + k, block=min(n, 1 << 16)) for k in range(s)])
    if packed:
        host = np.stack([synth.pack10(r) for r in xs])
        want = ddc.run_tensor(torch.from_numpy(host).cuda(), 100e6, packed=True).cpu().numpy()
        run = lambda: ddc.run_batch_packed(host, 100e6)   # noqa: E731
    else:
        host = xs.astype(np.float32)
        want = ddc.run_tensor(torch.from_numpy(host).cuda(), 100e6).cpu().numpy()
        run = lambda: ddc.run_batch(host, 100e6)   # noqa: E731
    scale = np.abs(want).max()
    for mode, chunk in ((0, 1 << 24), (0, 3 * n + 5), (0, n), (1, 1 << 24), (1, n // 3), (0, n // 3)):
        ddc.set_option("host_chunk_mode", mode)
        ddc.set_option("chunk_samples", chunk)
        got = run()
        assert got.shape == want.shape and np.abs(got - want).max() <= TOL_MAX * scale, (mode, chunk, np.abs(got - want).max() / scale)


def test_tensor_engine_fuzz(tmp_path):
    """Seeded random (T, D, N, streams, fc) on packed input: filters shorter than the decimation, two taps, tap counts that are
    no multiple of anything, lengths from a single output to several tiles.  Guards the geometry seams of the engine (row
    width 64 / 128, K padding, ragged last tile, halo shorter than a row when T < D)."""
    from scipy import signal

    rng = np.random.default_rng(20261019)
    seen = set()
    for case in range(40):
        d = int(rng.choice([4, 8, 16, 32, 64]))
        t = int(rng.choice([2, 3, rng.integers(3, 40), rng.integers(40, 300), rng.integers(300, 1400), 64, 256, 1024]))   # a one-line CSV loads as a 0-d array, in the reference too
        streams = int(rng.integers(1, 5))
        n = max(t, int(rng.integers(1, 3000)) * 64)         # rows stay 16-byte aligned
        n = (n + 63) // 64 * 64
        fc = float(rng.choice([100e6, 53.5e6, 428e6, 1.0e6, 855e6]))
        tp = signal.firwin(t, min(0.8 / d, 0.99)) if t > 3 else np.ones(t)
        ddc = DigitalDownConverter(d, FS, _custom_taps(tmp_path, tp))
        xs = np.stack([synth.digitiser_stream(n, 2000 + case * 7 + s) for s in range(streams)])
        if case % 3 == 0:
            xs = _full_range_streams(n)[:1]
        y = _run_packed(ddc, xs, fc=fc).cpu().numpy()
        seen.add(ddc.last_variant.split(",K")[0])
        ref = np.stack([orc.ddc_reference(r.astype(np.float32), fc, tp, d, FS) for r in xs])
        assert y.shape == ref.shape, (case, d, t, n, y.shape, ref.shape)
        emax, el2 = rel_err(y, ref)
        k = 4 if t > 256 else 1
        assert emax <= k * TOL_MAX and el2 <= k * TOL_L2, (case, d, t, n, streams, fc, ddc.last_variant, emax, el2)
        ddc.close()
    print("engine geometries exercised:", sorted(seen))
    assert sum("tensor_fir" in v for v in seen) >= 6, seen
