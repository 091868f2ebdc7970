#!/usr/bin/env python
"""End-to-end host path against bare copies, all ranks of one box AT THE SAME TIME.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/e2e_probe.py

Every rank owns 16 streams x 2^24 float32 samples in pinned host memory (1 GiB in, 128 MiB out: one rank's share of BASELINE
configs[4] at 8 GPUs).  Every measurement starts from a barrier, so the ranks really compete for the host's memory and PCIe
uplinks (round 1's probe took a per-rank minimum over unsynchronised repeats, which is a best case, not a concurrent one):
  bare        one flat 1 GiB H2D + one flat 128 MiB D2H on two streams, nothing else
  bare16      the same bytes as 16 flat H2D copies of 64 MiB + 16 D2H of 8 MiB (the by-stream chunking, no kernels)
  e2e_time    ddcb200_run_host_f32, time chunks (2-D copies of 16 row pieces; option host_chunk_mode = 1: round 1's path)
  e2e_stream  ddcb200_run_host_f32, by-stream chunks (flat copies; the default)
  e2e_big     by-stream chunks of 4 streams (chunk_samples = 2^26)
each as `burst` (barrier, one step, per-rank time) and `sustained` (barrier, 5 steps back to back, per-rank time per step).
Rank 0 prints one table."""
import os
import sys
import tempfile
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dc_sand_b200 import DigitalDownConverter, _lib, taps  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if os.environ.get("E2E_DEVICES"):                       # placement experiment: local rank -> device index
        local = int(os.environ["E2E_DEVICES"].split(",")[local])
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    s, n = 16, 1 << 24
    ddc = DigitalDownConverter(16, 1712e6, taps.write_csv("ddc_coeff_107MHz.csv", tempfile.mkdtemp()), device=local)
    m = ddc.out_len(n)
    h_in = torch.empty((s, n), dtype=torch.float32, pin_memory=True)
    h_in.fill_(1.0)
    h_out = torch.empty((s, m), dtype=torch.complex64, pin_memory=True)
    h_out.fill_(0)
    d_in = torch.empty((s, n), dtype=torch.float32, device="cuda")
    d_out = torch.empty((s, m), dtype=torch.complex64, device="cuda")
    lib = _lib.load()
    hnd = ddc._get_handle()
    step = ddc.phase_step(n, 100e6)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def bare():
        with torch.cuda.stream(s_in):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s_out):
            h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()

    def bare16():
        for k in range(s):
            with torch.cuda.stream(s_in):
                d_in[k].copy_(h_in[k], non_blocking=True)
            with torch.cuda.stream(s_out):
                h_out[k].copy_(d_out[k], non_blocking=True)
        torch.cuda.synchronize()

    def e2e():
        _lib.check(lib.ddcb200_run_host_f32(hnd, h_in.data_ptr(), n, s, h_in.stride(0), step, 0, h_out.data_ptr(), m), "run_host")

    def e2e_mode(mode, chunk):
        def f():
            e2e()
        f.setup = lambda: (ddc.set_option("host_chunk_mode", mode), ddc.set_option("chunk_samples", chunk))
        return f

    cases = [("bare", bare), ("bare16", bare16), ("e2e_time", e2e_mode(1, 1 << 24)), ("e2e_stream", e2e_mode(0, 1 << 24)),
             ("e2e_big", e2e_mode(0, 1 << 26))]
    if os.environ.get("E2E_CASES"):
        cases = [c for c in cases if c[0] in os.environ["E2E_CASES"].split(",")]
    rows = []
    for name, fn in cases:
        if hasattr(fn, "setup"):
            fn.setup()
        fn()                                  # warm-up (allocations, first-touch)
        burst = []
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            fn()
            burst.append(time.perf_counter() - t0)
        barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        sus = (time.perf_counter() - t0) / 5
        rows.append((name, min(burst) * 1e3, sorted(burst)[1] * 1e3, sus * 1e3))
    if world > 1:
        allrows = [None] * world
        dist.all_gather_object(allrows, rows)
    else:
        allrows = [rows]
    if rank == 0:
        gib = s * n * 4 / 2**30
        print(f"# devices {os.environ.get('E2E_DEVICES', 'identity')}")
        print(f"# {world} rank(s), each {gib:.0f} GiB H2D + {s * m * 8 / 2**20:.0f} MiB D2H per step; ms per step per rank (burst = median of 3 "
              "barrier-started single steps, sustained = 5 back-to-back steps after one barrier)")
        for i, (name, *_rest) in enumerate(rows):
            b = [allrows[r][i][2] for r in range(world)]
            su = [allrows[r][i][3] for r in range(world)]
            agg = world * s * n * 4 / (max(su) * 1e-3) / 1e9
            print(f"{name:11s} burst  " + " ".join(f"{v:6.1f}" for v in b) + f"   max {max(b):6.1f}")
            print(f"{name:11s} sustnd " + " ".join(f"{v:6.1f}" for v in su) + f"   max {max(su):6.1f}  -> job H2D {agg:6.1f} GB/s, "
                  f"{world * s * n / (max(su) * 1e-3) / 1e9:6.1f} Gsamples/s")
    ddc.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
