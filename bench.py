#!/usr/bin/env python
"""Benchmark of the fused digital down-converter (BASELINE.json: "DDC input Gsamples/s and % of HBM roofline").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c5] [--no-extra]

One "step" = one pass of the hot path (NCO mix -> FIR -> decimate) over one batch of synthetic digitiser samples.
  N = 1  : BASELINE configs[1]  "single L-band stream (1712 MSPS), 2^28 samples, 256 taps, decimation 16" is the headline
           `value`; the same line carries the other configs measured in the same run under `extra`: `c3` (64 packed
           streams), `sweep` / `sweep_packed` (the 25 tap x decimation cells), `c5_g1` (the 128-stream workload of
           configs[4] on one GPU: the G = 1 point of the strong-scaling curve) and `e2e_run_api` (the drop-in
           DigitalDownConverter.run() on a pageable NumPy array, complex128 out).
  N > 1  : launched with torch.distributed.run, one rank per GPU; BASELINE configs[4]: the SAME 128 streams x 2^24 samples
           at every N, 128 / N per GPU (strong scaling), no collective on the data path; device-timed, max over ranks.
Inputs are generated in HBM by the library's own test-vector generator (ddcb200_cwg, digitiser model) and, for the e2e leg,
copied once to pinned host memory.  `value` is device-resident throughput (inputs already in HBM); `e2e` is the same metric
through the reference-facing host API (pinned host buffers, H2D and D2H inside the timed region).  Every rank checks two
512-output windows of what the timed launches produced against the float64 windowed oracle (`parity_max_err`, outside the
timed region).  `--impl reference` times the CPU restatement of the reference (oracle/ddc_oracle.py; the reference itself
cannot travel to the GPU box) on all host cores.  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 1712e6
FC = 100e6
T = 256
D = 16
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.4: CUDA-core FMA peak at clocks.max.sm


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload: str, variant: str):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel from the committed
    `ncu --set full` capture of this workload (profiles/ncu_traffic.json), with the capture it comes from; (None, None) if
    the capture is of another kernel."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(p) as f:
            rec = json.load(f).get(workload)
    except (OSError, ValueError):
        return None, None
    if not rec or not variant.startswith(rec.get("variant_prefix", "\0")):
        return None, None
    return rec.get("dram_bytes_per_launch"), rec.get("source")


class ClockSampler:
    """Samples SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def bind_to_gpu_numa(index: int):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned host buffers are allocated
    (first touch puts them on that node).  With one rank per GPU this keeps every rank's H2D / D2H traffic on its own
    memory controller; without it the end-to-end leg of an 8-rank run shares one node's bandwidth.  Best effort."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(index).pci_bus_id   # torch >= 2.x: "0000:1B:00.0"-like
    except Exception:
        bus = None
    try:
        if bus is None:
            import pynvml

            pynvml.nvmlInit()
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


# ----------------------------------------------------------------------------------------------------------------
# CPU arms
# ----------------------------------------------------------------------------------------------------------------
def _cpu_call(args):
    seed, n = args
    from dc_sand_b200 import synth, taps
    from oracle import ddc_oracle as orc

    x = synth.digitiser_stream(n, seed).astype(np.float32)
    tp = taps.coefficients("ddc_coeff_107MHz.csv")
    t0 = time.perf_counter()
    y = orc.ddc_reference(x, FC, tp, D, FS, faithful_noise=True)
    return time.perf_counter() - t0, len(y)


def _cpu_fir_only(n):
    """Time of the FIR + decimation stages alone (ddc.py:98,119) on an already mixed complex64 array, one thread."""
    from scipy import signal

    from dc_sand_b200 import synth, taps
    from oracle import ddc_oracle as orc

    x = synth.digitiser_stream(n, 7).astype(np.float32)
    tp = taps.coefficients("ddc_coeff_107MHz.csv")
    mix = x * orc.nco(n, FC, FS)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        y = (signal.convolve(mix, tp, mode="valid") / sum(tp))[0::D]
        best = min(best, time.perf_counter() - t0)
    return best, len(y)


def cpu_baseline_single_core(n_calls=6, n=1 << 20):
    """The faithful restatement of the reference's run() (incl. its zero-scaled noise draw), one thread."""
    _cpu_call((1, n))  # warm-up
    t = [_cpu_call((2 + i, n))[0] for i in range(n_calls)]
    best = min(t)
    fir_s, _ = _cpu_fir_only(n)
    return {
        "fir_only_value": n / fir_s / 1e9,   # the FIR + decimate stages without NCO / noise generation
        "value": n / best / 1e9,
        "unit": "Gsamples/s",
        "cores": 1,
        "kind": "port",
        "sample": f"best of {n_calls} calls of oracle.ddc_reference(faithful_noise=True) on 2^20 samples (config 1), "
                  f"mean {np.mean(t):.3f} s/call; the reference is single-threaded",
        "host_cores_available": os.cpu_count(),
    }


def run_reference_arm(args):
    """--impl reference: the CPU restatement on every host core (one 2^20-sample call per worker per step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n = 1 << 20
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for w in range(args.warmup):
            pool.map(_cpu_call, [(1000 + w * cores + i, n) for i in range(cores)])
        t0 = time.perf_counter()
        for k in range(args.steps):
            pool.map(_cpu_call, [(5000 + k * cores + i, n) for i in range(cores)])
        dt = time.perf_counter() - t0
    value = cores * n * args.steps / dt / 1e9
    line = {
        "impl": "reference",
        "metric": "ddc_input_gsamples_per_s",
        "value": value,
        "unit": "Gsamples/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64 (complex128 FIR, complex64 mixer), as the reference",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "taps": T, "decimation": D,
                   "sample_per_step": f"{cores} x 2^20-sample calls (one per host core)"},
        "cpu_baseline": {"value": value, "unit": "Gsamples/s", "cores": cores, "kind": "port",
                         "sample": f"{cores} workers x {args.steps} steps x 2^20 samples of oracle.ddc_reference("
                                   "faithful_noise=True); /root/reference is Python and does not exist on this box"},
        "e2e": {"value": value, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


TOTAL_STREAMS_C5 = 128          # BASELINE configs[4]: 64 antennas x 2 polarisations


def workload_name(args):
    if args.workload == "c3":
        return "64 streams x 2^24 packed 10-bit samples, 256 taps, decimation 16 (BASELINE configs[2])"
    if args.gpus == 1 and args.workload == "c2":
        return "single L-band stream (1712 MSPS), 2^28 float32 samples, 256 taps, decimation 16 (BASELINE configs[1])"
    per = TOTAL_STREAMS_C5 // max(args.gpus, 1)
    return (f"{TOTAL_STREAMS_C5} streams x 2^24 float32 samples sharded by stream over {args.gpus} GPU(s), {per} per GPU, "
            "256 taps, decimation 16 (BASELINE configs[4])")


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
def roofline_of(n_streams, n, m, t, d, packed, kern_s, variant, hbm_peak, peak_src, traffic=(None, None)):
    """algorithmic bytes = (B_in + 8 / D) per input sample, direct-form flops = 4 T per output; the fast-FIR
    kernels EXECUTE 3/4 of those multiplies (plus one packed subtraction per window sample, not counted).  `bound` is the
    floor that binds: max(bytes / HBM peak, executed flops / CUDA-core FP32 peak); `floor_frac` = t_floor / t."""
    b_in = 1.25 if packed else 4.0
    alg_bytes = n_streams * n * (b_in + 8.0 / d)
    flops = 4.0 * t * m * n_streams
    tensor = variant.startswith("tensor_fir")   # packed input on tcgen05: no FIR flops on the CUDA cores at all
    exec_flops = 0.0 if tensor else flops * (0.75 if "fast_fir" in variant else 1.0)
    tensor_flops = None
    if tensor:   # MACs the engine issues, zeros of the banded tap matrix included: 128 x N x K per tile of 128 rows
        import re

        g = {k: int(v) for k, v in re.findall(r"(ROW|N|K)(\d+)", variant)}
        tiles = -(-m // (128 * (g["ROW"] // d))) * n_streams
        tensor_flops = 2.0 * 128 * g["N"] * g["K"] * tiles
    t_hbm = alg_bytes / (hbm_peak * 1e9)
    t_fp32 = exec_flops / (FP32_PEAK_TFLOPS * 1e12)
    achieved = alg_bytes / kern_s / 1e9
    return {
        "bound": "hbm" if t_hbm >= t_fp32 else "fp32",
        "achieved": achieved,
        "peak": hbm_peak,
        "unit": "GB/s",
        "frac": achieved / hbm_peak,
        "traffic": traffic[0],
        "traffic_source": traffic[1],
        "peak_source": peak_src,
        "frac_of_nominal_8TBs": achieved / 8000.0,
        "floor_frac": max(t_hbm, t_fp32) / kern_s,
        "t_floor_hbm_ms": t_hbm * 1e3,
        "t_floor_fp32_executed_ms": t_fp32 * 1e3,
        "fp32_direct_form_tflops": flops / kern_s / 1e12,
        "fp32_executed_tflops": exec_flops / kern_s / 1e12,
        "fp32_executed_frac": exec_flops / kern_s / 1e12 / FP32_PEAK_TFLOPS,
        "fp32_peak_tflops": FP32_PEAK_TFLOPS,
        "kernel_ms_mean": kern_s * 1e3,
        "algorithmic_bytes_per_launch": alg_bytes,
        "flop_per_launch_direct_form": flops,
        "engine": "tcgen05 (fp16 samples x hi+lo fp16 taps, FP32 accumulate in TMEM)" if tensor else "CUDA cores (FP32 FFMA2)",
        "tensor_tflops_issued": None if tensor_flops is None else tensor_flops / kern_s / 1e12,
    }


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    if world > 1:   # fewer ranks than GPUs: deal the ranks over both host-bridge groups of the box (scheduler.py)
        from dc_sand_b200.scheduler import device_for_local_rank

        local = device_for_local_rank(local, int(os.environ.get("LOCAL_WORLD_SIZE", world)), torch.cuda.device_count(),
                                      os.environ.get("DDCB200_DEVICE_ORDER"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = bind_to_gpu_numa(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from dc_sand_b200 import DigitalDownConverter, _lib, cwg as dcwg, taps
    from dc_sand_b200.scheduler import shard_range
    from oracle import ddc_oracle as orc      # the checker of the parity spot checks below, never timed on this arm

    lib = _lib.load()
    tmp = tempfile.mkdtemp()
    csv107 = taps.write_csv("ddc_coeff_107MHz.csv", tmp)
    hbm_peak, peak_src = measured_peaks()
    stream = torch.cuda.Stream(device=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def gen_input(n_streams, n, packed, seed):
        """Synthetic digitiser streams built in HBM (ddcb200_cwg, digitiser model of SURVEY 8d: tone + Gaussian noise, rounded
        and clipped to 10 bits; Philox keyed by (seed, stream)), packed on the device when asked."""
        x = dcwg.generate_carrier_wave_gpu(100.0, FC + 3.3e6, FS, n, 40.0, False, seed=seed, device=local, n_streams=n_streams,
                                           digitise=True)
        if packed:
            xp = dcwg.pack10_gpu(x)
            del x
            return xp
        return x

    def time_device(ddc, d_in, d_out, packed, steps, warm, sampler=None):
        """`steps` launches queued back to back on one stream between two CUDA events, behind a ~2 ms device-side delay so
        that the host enqueues while the GPU waits (no launch latency inside the region).  Returns ms for all steps."""
        with torch.cuda.stream(stream):
            for _ in range(warm):
                ddc.run_tensor(d_in, FC, out=d_out, packed=packed)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            torch.cuda._sleep(4_000_000)
            e0.record(stream)
            for _ in range(steps):
                ddc.run_tensor(d_in, FC, out=d_out, packed=packed)
            e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def parity_window(ddc, d_in, d_out, packed, n, t, d, seed):
        """max |y - oracle| / max|y| over one 512-output window of two streams (head window of stream 0, a random window of
        another): the float64 windowed oracle on the samples read back from HBM.  Outside every timed region."""
        rng = np.random.default_rng(seed)
        m = d_out.shape[1]
        step = orc.phase_step_cycles(n, FC, FS)
        scale = float(d_out[:, : min(m, 1 << 16)].abs().max())
        worst = 0.0
        for s_i, m0 in ((0, 0), (int(rng.integers(0, d_in.shape[0])), int(rng.integers(0, max(m - 512, 1))) // 4 * 4)):
            cnt = min(512, m - m0)
            if packed:
                seg = orc.unpack10(d_in[s_i, m0 * d // 4 * 5: ((m0 + cnt - 1) * d + t + 3) // 4 * 5].cpu().numpy()).astype(np.float32)
            else:
                seg = d_in[s_i, m0 * d: (m0 + cnt - 1) * d + t].cpu().numpy()
            ref = orc.ddc_windowed_f64(seg, m0, cnt, step, ddc.ddc_filter_coeffs, d, x_base=m0 * d)
            worst = max(worst, float(np.abs(d_out[s_i, m0:m0 + cnt].cpu().numpy() - ref).max()) / scale)
        return worst

    # ---- headline workload ---------------------------------------------------------------------------------------
    packed = args.workload == "c3"
    if packed:
        n_streams, n = 64, 1 << 24
    elif args.gpus == 1 and args.workload == "c2":
        n_streams, n = 1, 1 << 28
    else:
        a, b = shard_range(TOTAL_STREAMS_C5, world, rank)     # strong scaling: the same 128 streams at every N
        n_streams, n = b - a, 1 << 24
    if args.samples:
        n = args.samples
    if args.streams:
        n_streams = args.streams
    total_streams = n_streams * world if (args.streams or packed or (args.gpus == 1 and args.workload == "c2")) else TOTAL_STREAMS_C5

    ddc = DigitalDownConverter(D, FS, csv107, device=local)
    m = ddc.out_len(n)
    d_in = gen_input(n_streams, n, packed, seed=1234 + 4096 * rank)
    d_out = torch.empty((n_streams, m), dtype=torch.complex64, device=dev)
    in_bytes = d_in.numel() * d_in.element_size()
    out_bytes = n_streams * m * 8

    time_device(ddc, d_in, d_out, packed, 1, max(args.warmup, 3))
    barrier()
    launches0 = ddc.launch_count
    with ClockSampler(local) as clk:
        total_ms = time_device(ddc, d_in, d_out, packed, args.steps, 0)
    launches = ddc.launch_count - launches0
    variant = ddc.last_variant
    barrier()
    parity_err = parity_window(ddc, d_in, d_out, packed, n, T, D, seed=rank)

    # ---- e2e: pinned host buffers through the C ABI (H2D + kernel + D2H per step) ---------------------------------
    # Several ranks of one box share its host links, and not evenly (profiles/r2_n8_placement.txt), while what bounds a
    # host-fed step is exactly that link: the end-to-end leg of the multi-GPU job therefore shards the SAME 128 streams by the
    # host-to-device rate each rank measures with all ranks copying at once (scheduler.weighted_shard_sizes; equal shards when
    # the rates agree within 15 %).  The device-resident leg above keeps the balanced shards: there the GPU is the bound.
    e2e_streams, e2e_sizes, probe_rates = n_streams, None, None
    if world > 1 and not packed and not args.streams:
        from dc_sand_b200.scheduler import weighted_shard_sizes

        forced = os.environ.get("DDCB200_E2E_WEIGHTS")          # "w0,w1,..." (testing); "equal" switches the weighting off
        if forced == "equal":
            probe_rates = [1.0] * world
        elif forced:
            probe_rates = [float(v) for v in forced.split(",")]
        else:
            pb = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
            db = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
            db.copy_(pb, non_blocking=True)
            torch.cuda.synchronize()
            barrier()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            for _ in range(3):
                db.copy_(pb, non_blocking=True)
            p1.record()
            torch.cuda.synchronize()
            mine = torch.tensor([3 * (256 << 20) / (p0.elapsed_time(p1) * 1e-3) / 1e9], dtype=torch.float64, device=dev)
            allr = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            probe_rates = [float(v.item()) for v in allr]
            del pb, db
        med = float(np.median(probe_rates))
        probe_rates = [min(max(v, 0.5 * med), 2.0 * med) for v in probe_rates]   # one wild reading must not empty or flood a rank
        e2e_sizes = weighted_shard_sizes(TOTAL_STREAMS_C5, probe_rates)
        e2e_streams = e2e_sizes[rank]
    h_in = torch.empty((e2e_streams, d_in.shape[1]), dtype=d_in.dtype, pin_memory=True)
    for r0 in range(0, e2e_streams, n_streams):                  # a rank with more streams than its device-resident shard repeats rows
        cnt = min(n_streams, e2e_streams - r0)
        h_in[r0:r0 + cnt].copy_(d_in[:cnt])
    h_out = torch.empty((e2e_streams, m), dtype=torch.complex64, pin_memory=True)
    torch.cuda.synchronize()
    step = ddc.phase_step(n, FC)
    fn = lib.ddcb200_run_host_packed10 if packed else lib.ddcb200_run_host_f32
    h = ddc._get_handle()

    def e2e_pass():
        _lib.check(fn(h, h_in.data_ptr(), n, e2e_streams, h_in.stride(0), step, 0, h_out.data_ptr(), m), "run_host")

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_pass()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_pass()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    # (tolerance, not bit equality: the host path runs time chunks, which may pair outputs differently in the fast FIR)
    rows = min(e2e_streams, n_streams)
    a_dev, a_host = d_out[:rows, : min(m, 4096)].cpu(), h_out[:rows, : min(m, 4096)]
    same = bool((a_dev - a_host).abs().max() <= 1e-5 * a_dev.abs().max())
    e2e_in_bytes = h_in.numel() * h_in.element_size()
    e2e_out_bytes = e2e_streams * m * 8
    del h_in, h_out

    # ---- optional final gather of the outputs over NCCL (outside the timed region; the hot path has no collective) ---
    gather_ms = None
    if world > 1:
        from dc_sand_b200.scheduler import ShardedDDC

        sh = ShardedDDC(TOTAL_STREAMS_C5 if not args.streams else world * n_streams, rank, world, ddc=ddc, center_freq=FC)
        torch.cuda.synchronize()
        dist.barrier()
        g0 = time.perf_counter()
        full = sh.gather(d_out, dst=0)
        torch.cuda.synchronize()
        gather_ms = (time.perf_counter() - g0) * 1e3
        if rank == 0:
            assert full.shape == (sh.n_streams, m)
            assert torch.equal(full[:n_streams], d_out)
        del full

    # ---- max over ranks -------------------------------------------------------------------------------------------
    if os.environ.get("DDCB200_BENCH_VERBOSE"):
        print(f"[rank {rank}] device {total_ms / args.steps:.4f} ms/step, e2e {e2e_s * 1e3:.2f} ms/step, parity {parity_err:.2e}",
              file=sys.stderr, flush=True)
    stats = torch.tensor([total_ms, e2e_s, parity_err], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    total_ms, e2e_s, parity_err = [float(v) for v in stats.cpu()]

    extra = {}
    if rank == 0 and world == 1 and args.workload == "c2" and not args.no_extra and not args.samples and not args.streams:
        del d_in, d_out
        torch.cuda.empty_cache()
        extra = run_extras(args, torch, dev, local, gen_input, time_device, parity_window, hbm_peak, peak_src, csv107, tmp)

    if rank == 0:
        samples_per_step = total_streams * n
        value = samples_per_step * args.steps / (total_ms * 1e-3) / 1e9
        # average launch duration over the timed region (the region holds nothing but these launches, back to back)
        kern_s = total_ms * 1e-3 / max(int(launches), 1)
        wl = args.workload if (args.gpus == 1 or args.workload != "c2") else "c5"
        line = {
            "metric": "ddc_input_gsamples_per_s",
            "value": value,
            "unit": "Gsamples/s",
            "n_gpus": args.gpus,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps,
            "higher_is_better": True,
            # configs[4] is a fixed job (128 streams) cut over N GPUs; its G = 1 point is extra.c5_g1 of the N = 1 line (the N = 1
            # headline is configs[1], one 2^28-sample stream: the same samples per GPU as a rank of the N = 8 run)
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": workload_name(args),
                "total_streams": total_streams,
                "streams_per_gpu": n_streams,
                "samples_per_stream": n,
                "taps": T,
                "decimation": D,
                "input": "packed10" if packed else "float32",
                "l2": f"input per launch {in_bytes / 2**20:.0f} MiB >> 126 MB L2 (no flush needed)",
                "kernel": variant,
            },
            "roofline": roofline_of(n_streams, n, m, T, D, packed, kern_s, variant, hbm_peak, peak_src, ncu_traffic(wl, variant)),
            "e2e": {
                "value": total_streams * n / e2e_s / 1e9,
                "unit": "Gsamples/s",
                "h2d_bytes_per_step": e2e_in_bytes,
                "d2h_bytes_per_step": e2e_out_bytes,
                "ms_per_step": e2e_s * 1e3,
                "steps": e2e_steps,
                "api": "ddcb200_run_host_* (C ABI, pinned host buffers, complex64 out)",
                "matches_device_path": same,
                # multi-GPU: streams per rank of this leg and the probe behind them (bytes above are rank 0's)
                "streams_per_rank": e2e_sizes,
                "host_link_probe_gbps": None if probe_rates is None else [round(v, 1) for v in probe_rates],
                "partition": None if e2e_sizes is None else "streams sharded in proportion to each rank's host-to-device rate with all ranks "
                                                            "copying at once (dc_sand_b200.scheduler.weighted_shard_sizes; balanced within 15 %)",
            },
            "parity_max_err": parity_err,
            "parity_check": "max over ranks of |y - float64 windowed oracle| / max|y| on two 512-output windows per rank (tolerance 1e-5)",
            "gpu_launches": int(launches),
            "final_gather_ms_untimed": gather_ms,
            "numa_node_rank0": numa_node,
            "device_rank0": local,
            "device_placement": "ranks dealt alternately to the two halves of the box's devices when fewer ranks than devices "
                                "(dc_sand_b200.scheduler.device_for_local_rank; profiles/r2_n8_placement.txt)",
            "clocks": clk.summary(),
        }
        if extra:
            line["extra"] = extra
        if args.gpus == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_single_core()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_extras(args, torch, dev, local, gen_input, time_device, parity_window, hbm_peak, peak_src, csv107, tmp):
    """The other BASELINE configs on the same box, in the same run (one GPU): configs[2] (64 packed streams), configs[3] (the
    25-cell tap / decimation sweep), configs[4] at G = 1 (the 128-stream workload on one GPU: the denominator of the strong-
    scaling efficiency) and the drop-in DigitalDownConverter.run() on a pageable NumPy array."""
    from scipy import signal

    from dc_sand_b200 import DigitalDownConverter, _lib

    lib = _lib.load()
    extra = {}

    # ---- configs[2]: 64 streams x 2^24 packed 10-bit samples ---------------------------------------------------------
    def _c3():
        s3, n3 = 64, 1 << 24
        ddc = DigitalDownConverter(D, FS, csv107, device=local)
        m3 = ddc.out_len(n3)
        x3 = gen_input(s3, n3, True, seed=777)
        y3 = torch.empty((s3, m3), dtype=torch.complex64, device=dev)
        steps3 = 10
        ms3 = time_device(ddc, x3, y3, True, steps3, 3) / steps3
        v3 = ddc.last_variant
        err3 = parity_window(ddc, x3, y3, True, n3, T, D, seed=3)
        ddc.set_option("packed_engine", 0)          # the same workload on the CUDA-core fused-unpack kernel, for comparison
        ms3c = time_device(ddc, x3, y3, True, steps3, 3) / steps3
        v3c = ddc.last_variant
        err3c = parity_window(ddc, x3, y3, True, n3, T, D, seed=3)
        ddc.set_option("packed_engine", 1)
        h_in = torch.empty(x3.shape, dtype=torch.uint8, pin_memory=True)
        h_in.copy_(x3)
        h_out = torch.empty((s3, m3), dtype=torch.complex64, pin_memory=True)
        torch.cuda.synchronize()
        step3 = ddc.phase_step(n3, FC)
        hnd = ddc._get_handle()
        e2e_t = []
        for i in range(3):
            t0 = time.perf_counter()
            _lib.check(lib.ddcb200_run_host_packed10(hnd, h_in.data_ptr(), n3, s3, h_in.stride(0), step3, 0, h_out.data_ptr(), m3))
            e2e_t.append(time.perf_counter() - t0)
        e2e3 = min(e2e_t[1:])
        extra["c3"] = {
            "workload": "64 streams x 2^24 packed 10-bit samples, 256 taps, decimation 16 (BASELINE configs[2])",
            "ms": ms3, "gsamples_per_s": s3 * n3 / ms3 / 1e6, "steps": steps3, "kernel": v3, "parity_max_err": err3,
            "roofline": roofline_of(s3, n3, m3, T, D, True, ms3 * 1e-3, v3, hbm_peak, peak_src, ncu_traffic("c3", v3)),
            "e2e": {"value": s3 * n3 / e2e3 / 1e9, "unit": "Gsamples/s", "ms_per_step": e2e3 * 1e3,
                    "h2d_bytes_per_step": x3.numel(), "d2h_bytes_per_step": s3 * m3 * 8},
            "cuda_cores": {"ms": ms3c, "gsamples_per_s": s3 * n3 / ms3c / 1e6, "kernel": v3c, "parity_max_err": err3c,
                           "hbm_frac": roofline_of(s3, n3, m3, T, D, True, ms3c * 1e-3, v3c, hbm_peak, peak_src)["frac"],
                           "note": "option packed_engine = 0: the round-1 fused-unpack kernel"},
        }
        del x3, y3, h_in, h_out
        ddc.close()
        torch.cuda.empty_cache()

    # ---- configs[3]: taps 64 .. 1024 x decimation 4 .. 64, N = 2^26 float32 (256 MiB per launch > L2) ------------------
    def _sweep():
        n4 = 1 << 26
        x4 = gen_input(1, n4, False, seed=4242)
        x4p = None
        sweep, sweep_packed = [], []
        for t in (64, 128, 256, 512, 1024):
            for d in (4, 8, 16, 32, 64):
                if t == 256:
                    csv = csv107
                else:
                    csv = os.path.join(tmp, f"firwin_{t}_{d}.csv")
                    np.savetxt(csv, signal.firwin(t, 0.8 / d), fmt="%.18e")
                c = DigitalDownConverter(d, FS, csv, device=local)
                m4 = c.out_len(n4)
                y4 = torch.empty((1, m4), dtype=torch.complex64, device=dev)
                for pk, dst in ((False, sweep), (True, sweep_packed)):
                    if pk and x4p is None:
                        from dc_sand_b200 import cwg as dcwg

                        x4p = dcwg.pack10_gpu(x4)
                    xin = x4p if pk else x4
                    ms = time_device(c, xin, y4, pk, 10, 3) / 10
                    var = c.last_variant
                    r = roofline_of(1, n4, m4, t, d, pk, ms * 1e-3, var, hbm_peak, peak_src)
                    dst.append({"taps": t, "decimation": d, "ms": ms, "gsamples_per_s": n4 / ms / 1e6, "kernel": var.split("<")[0],
                                "bound": r["bound"], "hbm_frac": r["frac"], "fp32_executed_frac": r["fp32_executed_frac"],
                                "floor_frac": r["floor_frac"],
                                "parity_max_err": parity_window(c, xin, y4, pk, n4, t, d, seed=t + d)})
                    if pk:   # the same cell on the CUDA-core kernels (option packed_engine = 0)
                        c.set_option("packed_engine", 0)
                        msc = time_device(c, xin, y4, pk, 10, 3) / 10
                        dst[-1].update({"ms_cuda_cores": msc, "kernel_cuda_cores": c.last_variant.split("<")[0]})
                        c.set_option("packed_engine", 1)
                del y4
                c.close()
        extra["sweep"] = {"workload": "1 stream x 2^26 float32 samples, firwin(T, 0.8 / D) taps (the shipped 107 MHz filter at T = 256); "
                                      "3 warm-up + 10 launches per cell (BASELINE configs[3])", "cells": sweep}
        extra["sweep_packed"] = {"workload": "the same cells on the packed 10-bit form of the same samples: default engine (tcgen05 tensor "
                                             "cores, unpack fused in every cell) and, as ms_cuda_cores, the CUDA-core kernels", "cells": sweep_packed}
        del x4, x4p
        torch.cuda.empty_cache()

    # ---- DigitalDownConverter.run(): the reference's own call, pageable float32 NumPy in, complex128 out (2^26 samples) ----
    def _run_api():
        n4 = 1 << 26
        x_np = gen_input(1, n4, False, seed=4242)[0].cpu().numpy()   # the sweep's samples
        ddc = DigitalDownConverter(D, FS, csv107, device=local)
        ddc.run(x_np[: 1 << 22], FC)
        tr = []
        for i in range(3):
            t0 = time.perf_counter()
            y = ddc.run(x_np, FC)
            tr.append(time.perf_counter() - t0)
        best = min(tr)
        extra["e2e_run_api"] = {
            "api": "DigitalDownConverter.run(np.ndarray float32 [2^26] pageable, 100e6) -> complex128 (feng/ddc/src/ddc.py:121)",
            "value": n4 / best / 1e9, "unit": "Gsamples/s", "ms_per_call": best * 1e3, "calls": 3,
            "h2d_bytes_per_step": x_np.nbytes, "d2h_bytes_per_step": int(y.shape[0]) * 8, "out_dtype": str(y.dtype),
        }
        del x_np, y
        ddc.close()

    # ---- configs[4] at G = 1: the 128-stream workload on ONE GPU (8 GiB in, 1 GiB out) -------------------------------------
    def _c5_g1():
        s5, n5 = TOTAL_STREAMS_C5, 1 << 24
        ddc = DigitalDownConverter(D, FS, csv107, device=local)
        m5 = ddc.out_len(n5)
        x5 = gen_input(s5, n5, False, seed=1234)
        y5 = torch.empty((s5, m5), dtype=torch.complex64, device=dev)
        steps5 = 5
        ms5 = time_device(ddc, x5, y5, False, steps5, 3) / steps5
        v5 = ddc.last_variant
        extra["c5_g1"] = {
            "workload": "128 streams x 2^24 float32 samples on ONE GPU (BASELINE configs[4] at G = 1: strong-scaling denominator)",
            "ms": ms5, "gsamples_per_s": s5 * n5 / ms5 / 1e6, "steps": steps5, "kernel": v5,
            "parity_max_err": parity_window(ddc, x5, y5, False, n5, T, D, seed=5),
            "roofline": roofline_of(s5, n5, m5, T, D, False, ms5 * 1e-3, v5, hbm_peak, peak_src),
        }
        del x5, y5
        ddc.close()
        torch.cuda.empty_cache()
    # every section on its own: a failure in one of them (say, no room for the 9 GiB of configs[4] on a shared device) must not
    # cost the headline line or the other sections; it is reported in place of the section's numbers
    for keys, section in ((("c3",), _c3), (("sweep", "sweep_packed"), _sweep), (("e2e_run_api",), _run_api), (("c5_g1",), _c5_g1)):
        try:
            section()
        except Exception as e:   # noqa: BLE001 -- reported, not swallowed
            for k in keys:
                extra.setdefault(k, {"error": f"{type(e).__name__}: {e}"})
            torch.cuda.empty_cache()
    return extra


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c5"])
    ap.add_argument("--samples", type=int, default=0, help="override samples per stream")
    ap.add_argument("--streams", type=int, default=0, help="override streams per GPU")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="N = 1 only: skip the other BASELINE configs (extra.*)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
