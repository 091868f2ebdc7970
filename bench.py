#!/usr/bin/env python
"""Benchmark of the fused digital down-converter (BASELINE.json: "DDC input Gsamples/s and % of HBM roofline").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c5]     (sweep: tools/sweep.py)

One "step" = one pass of the hot path (NCO mix -> FIR -> decimate) over one batch of synthetic digitiser samples.
  N = 1  : BASELINE configs[1]  "single L-band stream (1712 MSPS), 2^28 samples, 256 taps, decimation 16".
  N > 1  : launched with torch.distributed.run, one rank per GPU; every rank processes 16 streams x 2^24 samples
           (= 2^28 samples per GPU, so N = 8 is BASELINE configs[4]: 128 streams sharded by stream); weak scaling,
           no collective on the data path; device-timed, max over ranks.
`value` is device-resident throughput (inputs already in HBM); `e2e` is the same metric through the reference-
facing host API (pinned host buffers, H2D and D2H inside the timed region).  `--impl reference` times the CPU
restatement of the reference (oracle/ddc_oracle.py, the reference itself cannot travel to the GPU box) on all
host cores.  Prints ONE JSON line.
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
    `ncu --set full` capture of this workload (profiles/ncu_traffic.json); None if the capture is of another kernel."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(p) as f:
            rec = json.load(f).get(workload)
    except (OSError, ValueError):
        return None
    if not rec or not variant.startswith(rec.get("variant_prefix", "\0")):
        return None
    return rec.get("dram_bytes_per_launch")


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


def make_input(n_streams: int, n: int, rank: int, packed: bool):
    """Synthetic tone + noise, 10-bit quantised. Large arrays reuse a 2^22-sample block per stream (rolled)."""
    from dc_sand_b200 import synth

    rows = []
    for s in range(n_streams):
        seed = 1234 + rank * 4096 + s
        v = synth.digitiser_stream_fast(n, seed, block=min(n, 1 << 22)) if n > (1 << 22) else synth.digitiser_stream(n, seed)
        rows.append(synth.pack10(v) if packed else v.astype(np.float32))
    return np.stack(rows)


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


def workload_name(args):
    if args.workload == "c3":
        return "64 streams x 2^24 packed 10-bit samples, 256 taps, decimation 16 (BASELINE configs[2])"
    if args.gpus == 1 and args.workload == "c2":
        return "single L-band stream (1712 MSPS), 2^28 float32 samples, 256 taps, decimation 16 (BASELINE configs[1])"
    return (f"{16 * args.gpus} streams x 2^24 float32 samples sharded by stream over {args.gpus} GPU(s), 16 per GPU, "
            "256 taps, decimation 16 (BASELINE configs[4] at 8 GPUs)")


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
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
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = bind_to_gpu_numa(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from dc_sand_b200 import DigitalDownConverter, _lib, taps

    packed = args.workload == "c3"
    if packed:
        n_streams, n = 64, 1 << 24
    elif args.gpus == 1 and args.workload == "c2":
        n_streams, n = 1, 1 << 28
    else:
        n_streams, n = 16, 1 << 24
    if args.samples:
        n = args.samples
    if args.streams:
        n_streams = args.streams

    tmp = tempfile.mkdtemp()
    ddc = DigitalDownConverter(D, FS, taps.write_csv("ddc_coeff_107MHz.csv", tmp), device=local)
    m = ddc.out_len(n)
    lib = _lib.load()

    # ---- inputs: pinned host buffers (e2e) and a device-resident copy (value) -----------------------------------
    x_np = make_input(n_streams, n, rank, packed)
    h_in = torch.from_numpy(x_np).pin_memory()
    h_out = torch.empty((n_streams, m), dtype=torch.complex64).pin_memory()
    d_in = h_in.to(dev, non_blocking=True)
    d_out = torch.empty((n_streams, m), dtype=torch.complex64, device=dev)
    torch.cuda.synchronize()
    in_bytes = x_np.nbytes
    out_bytes = n_streams * m * 8
    del x_np

    stream = torch.cuda.Stream(device=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def device_pass(steps, per_launch_events):
        evs = []
        with torch.cuda.stream(stream):
            for _ in range(steps):
                if per_launch_events:
                    e0 = torch.cuda.Event(enable_timing=True)
                    e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                ddc.run_tensor(d_in, FC, out=d_out, packed=packed)
                if per_launch_events:
                    e1.record(stream)
                    evs.append((e0, e1))
        return evs

    # warm-up
    device_pass(max(args.warmup, 3), False)
    torch.cuda.synchronize()
    barrier()
    launches0 = ddc.launch_count
    # Timed region: K steps queued back to back on one stream between two CUDA events (an event pair around every launch
    # would put two timestamp packets between consecutive kernels and measure those as well).
    with ClockSampler(local) as clk:
        torch.cuda.synchronize()
        t_start = torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            # a ~2 ms device-side delay in front of the start event: the host enqueues the K launches while it runs, so the
            # timed region holds K steps back to back and no host launch latency (about 0.1 ms before the first kernel)
            torch.cuda._sleep(4_000_000)
            t_start.record(stream)
        device_pass(args.steps, False)
        with torch.cuda.stream(stream):
            t_end.record(stream)
        torch.cuda.synchronize()
    launches = ddc.launch_count - launches0
    total_ms = t_start.elapsed_time(t_end)
    barrier()
    # diagnostic, outside the timed region: the same launches with an event pair around each one
    evs = device_pass(min(args.steps, 20), True)
    torch.cuda.synchronize()
    kern_ms = [a.elapsed_time(b) for a, b in evs]
    variant = ddc.last_variant

    # ---- e2e: host buffers through the C ABI (H2D + kernel + D2H per step) ---------------------------------------
    step = ddc.phase_step(n, FC)
    fn = lib.ddcb200_run_host_packed10 if packed else lib.ddcb200_run_host_f32
    stride_in = h_in.stride(0)
    h = ddc._get_handle()

    def e2e_pass():
        _lib.check(fn(h, h_in.data_ptr(), n, n_streams, stride_in, step, 0, h_out.data_ptr(), m), "run_host")

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_pass()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_pass()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    # parity spot check of what the timed calls produced (device-resident result vs host-path result)
    # (tolerance, not bit equality: the host path runs time chunks, which may pair outputs differently in the fast FIR)
    a_dev, a_host = d_out[:, : min(m, 4096)].cpu(), h_out[:, : min(m, 4096)]
    same = bool((a_dev - a_host).abs().max() <= 1e-5 * a_dev.abs().max())

    # ---- optional final gather of the outputs over NCCL (outside the timed region; the hot path has no collective) ---
    gather_ms = None
    if world > 1:
        from dc_sand_b200.scheduler import ShardedDDC

        sh = ShardedDDC(world * n_streams, rank, world, ddc=ddc, center_freq=FC)
        torch.cuda.synchronize()
        dist.barrier()
        g0 = time.perf_counter()
        full = sh.gather(d_out, dst=0)
        torch.cuda.synchronize()
        gather_ms = (time.perf_counter() - g0) * 1e3
        if rank == 0:
            assert full.shape == (world * n_streams, m)
            assert torch.equal(full[:n_streams], d_out)
        del full

    # ---- max over ranks -------------------------------------------------------------------------------------------
    if os.environ.get("DDCB200_BENCH_VERBOSE"):
        print(f"[rank {rank}] device {total_ms / args.steps:.4f} ms/step, e2e {e2e_s * 1e3:.2f} ms/step", file=sys.stderr, flush=True)
    stats = torch.tensor([total_ms, e2e_s, float(np.median(kern_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    total_ms, e2e_s, kern_med_ms = [float(v) for v in stats.cpu()]

    if rank == 0:
        samples_per_step = world * n_streams * n
        value = samples_per_step * args.steps / (total_ms * 1e-3) / 1e9
        hbm_peak, peak_src = measured_peaks()
        b_in = 1.25 if packed else 4.0
        alg_bytes = n_streams * n * (b_in + 8.0 / D)            # per launch, per GPU
        flops = 4.0 * T * m * n_streams                         # 2T real FMAs per output
        # average launch duration over the timed region (the region holds nothing but these launches, back to back)
        kern_s = total_ms * 1e-3 / max(int(launches), 1)
        achieved = alg_bytes / kern_s / 1e9
        line = {
            "metric": "ddc_input_gsamples_per_s",
            "value": value,
            "unit": "Gsamples/s",
            "n_gpus": args.gpus,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": workload_name(args),
                "streams_per_gpu": n_streams,
                "samples_per_stream": n,
                "taps": T,
                "decimation": D,
                "input": "packed10" if packed else "float32",
                "l2": f"input per launch {in_bytes / 2**20:.0f} MiB >> 126 MB L2 (no flush needed)",
                "kernel": variant,
            },
            "roofline": {
                "bound": "hbm",
                "achieved": achieved,
                "peak": hbm_peak,
                "unit": "GB/s",
                "frac": achieved / hbm_peak,
                "traffic": ncu_traffic(args.workload if (args.gpus == 1 or args.workload != "c2") else "c5", variant),
                "peak_source": peak_src,
                "frac_of_nominal_8TBs": achieved / 8000.0,
                "fp32_tflops": flops / kern_s / 1e12,
                "fp32_frac_of_74.4": flops / kern_s / 1e12 / FP32_PEAK_TFLOPS,
                "kernel_ms_mean": kern_s * 1e3,
                "kernel_ms_isolated_median": kern_med_ms,
                "algorithmic_bytes_per_launch": alg_bytes,
                "flop_per_launch": flops,
            },
            "e2e": {
                "value": world * n_streams * n / e2e_s / 1e9,
                "unit": "Gsamples/s",
                "h2d_bytes_per_step": in_bytes,
                "d2h_bytes_per_step": out_bytes,
                "ms_per_step": e2e_s * 1e3,
                "steps": e2e_steps,
                "matches_device_path": same,
            },
            "gpu_launches": int(launches),
            "final_gather_ms_untimed": gather_ms,
            "numa_node_rank0": numa_node,
            "clocks": clk.summary(),
        }
        if args.gpus == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_single_core()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
