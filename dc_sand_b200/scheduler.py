"""Stream scheduler: shards independent antenna-polarisation streams across the GPUs of one box.

Every stream of the reference operator is an independent 1-D problem (`DigitalDownConverter.run` is strictly 1-D,
/root/reference/feng/ddc/src/ddc.py:121-188), so the partition is by stream, there is NO collective on the data path
and no halo exchange.  Two ways to drive it:

  * one process per GPU under torch.distributed (`ShardedDDC`): rank r owns streams shard_range(S, world, r); an
    optional final gather of the [streams, M] complex64 outputs runs over NCCL (NVLink / NVSwitch) outside the hot path;
  * one process, several devices (`MultiDeviceDDC`): one native handle per device driven from Python threads (ctypes
    releases the GIL while the C ABI runs).

Only the bookkeeping lives here; all arithmetic happens in libddcb200.so.
"""
from __future__ import annotations

import threading
from typing import Callable, Sequence

import numpy as np


def shard_range(n_streams: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced block of streams for `rank`: the first n % world ranks get one extra stream."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size: {rank}/{world_size}")
    if n_streams < 0:
        raise ValueError("n_streams must be >= 0")
    base, extra = divmod(n_streams, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def device_for_local_rank(local_rank: int, local_world: int, n_devices: int, order: str | None = None) -> int:
    """CUDA device of a rank when FEWER ranks than devices share a box.

    The GPUs of an 8 x B200 box hang off two groups of host bridges (devices 0 .. 3 and 4 .. 7), and the host link is what
    bounds the end-to-end path: four ranks on devices 0 .. 3 get 100 GB/s between them, on 0, 4, 1, 5 they get 164 GB/s
    (profiles/r2_n8_placement.txt).  So ranks are dealt alternately to the two halves: 0, n/2, 1, n/2 + 1, ...  With as many
    ranks as devices it is the identity.  `order` ("4,5,6,7,0,1,2,3", or the environment variable DDCB200_DEVICE_ORDER in
    bench.py) overrides the rule."""
    if not (0 <= local_rank < local_world):
        raise ValueError(f"bad local rank {local_rank} of {local_world}")
    if order:
        devs = [int(v) for v in order.split(",")]
        if len(set(devs)) != len(devs) or len(devs) < local_world or any(not (0 <= d < n_devices) for d in devs):
            raise ValueError(f"device order {order!r} does not name {local_world} distinct devices below {n_devices}")
        return devs[local_rank]
    if local_world >= n_devices or n_devices < 2:
        if local_rank >= max(n_devices, 1):
            raise ValueError(f"local rank {local_rank} has no device ({n_devices} visible)")
        return local_rank
    half = (n_devices + 1) // 2
    dealt = [d for pair in zip(range(half), range(half, 2 * half)) for d in pair if d < n_devices]
    return dealt[local_rank]


def shard_sizes(n_streams: int, world_size: int) -> list[int]:
    return [b - a for a, b in (shard_range(n_streams, world_size, r) for r in range(world_size))]


def weighted_shard_sizes(n_streams: int, weights: Sequence[float], tolerance: float = 0.15) -> list[int]:
    """Streams per rank in proportion to `weights` (largest-remainder apportionment, every rank keeps at least one stream
    while there are enough).

    For the END-TO-END path of a box whose GPUs do not share the host evenly: streams are independent, so the partition is
    free, and what bounds a host-fed job is each rank's share of the host links, not its GPU (devices 0 .. 3 of this pool's
    8 x B200 boxes get ~25 GB/s each when all eight copy at once, devices 4 .. 7 ~47 GB/s: profiles/r2_n8_placement.txt).
    `weights` are the per-rank host-to-device rates measured with all ranks copying at once (bench.py does that in front of
    its end-to-end leg).  Weights within `tolerance` of each other (max / min - 1) give the balanced shard_sizes(): a noisy
    probe must not unbalance a symmetric box."""
    w = [float(v) for v in weights]
    world = len(w)
    if world == 0 or any(not (v > 0.0) or v != v or v == float("inf") for v in w):
        raise ValueError(f"weights must be positive and finite: {weights!r}")
    if n_streams < 0:
        raise ValueError("n_streams must be >= 0")
    if max(w) / min(w) - 1.0 <= tolerance or n_streams < world:
        return shard_sizes(n_streams, world)
    total = sum(w)
    quota = [n_streams * v / total for v in w]
    sizes = [max(1, int(q)) for q in quota]
    # hand out (or take back) the difference by largest (smallest) remainder; never below one stream
    while sum(sizes) < n_streams:
        r = max(range(world), key=lambda i: quota[i] - sizes[i])
        sizes[r] += 1
    while sum(sizes) > n_streams:
        r = min((i for i in range(world) if sizes[i] > 1), key=lambda i: quota[i] - sizes[i])
        sizes[r] -= 1
    return sizes


def weighted_shard_range(n_streams: int, weights: Sequence[float], rank: int, tolerance: float = 0.15) -> tuple[int, int]:
    """Contiguous block of streams of `rank` under weighted_shard_sizes()."""
    sizes = weighted_shard_sizes(n_streams, weights, tolerance)
    if not (0 <= rank < len(sizes)):
        raise ValueError(f"bad rank {rank} of {len(sizes)}")
    start = sum(sizes[:rank])
    return start, start + sizes[rank]


class ShardedDDC:
    """One rank of a stream-sharded down-converter (one process per GPU, torch.distributed initialised by the caller).

    `compute` maps a [s_local, N] tensor to a [s_local, M] complex64 tensor on the same device; by default it is
    `DigitalDownConverter.run_tensor` of this rank's device.  It is injectable so that the sharding / gather logic can
    be tested on CPU with the gloo backend (tests/test_scheduler.py) -- the product path always uses the CUDA operator.
    """

    def __init__(self, n_streams: int, rank: int, world_size: int, compute: Callable | None = None, ddc=None,
                 center_freq: float | None = None):
        self.n_streams = int(n_streams)
        self.rank, self.world_size = int(rank), int(world_size)
        self.start, self.stop = shard_range(self.n_streams, self.world_size, self.rank)
        if compute is None:
            if ddc is None or center_freq is None:
                raise ValueError("pass either compute= or ddc= and center_freq=")
            compute = lambda x: ddc.run_tensor(x, center_freq)  # noqa: E731
        self._compute = compute

    @property
    def local_streams(self) -> range:
        return range(self.start, self.stop)

    def run_local(self, x_local):
        """Down-convert this rank's streams. No communication."""
        if x_local.shape[0] != self.stop - self.start:
            raise ValueError(f"rank {self.rank} expects {self.stop - self.start} streams, got {x_local.shape[0]}")
        return self._compute(x_local)

    def gather(self, y_local, dst: int | None = None):
        """Optional final gather of the outputs, stream order preserved.

        dst=None -> every rank gets the full [streams, M] tensor (all_gather); dst=r -> only rank r (others get None).
        Shards may differ by one stream; they are padded to the largest shard for the collective and trimmed after.
        """
        import torch
        import torch.distributed as dist

        sizes = shard_sizes(self.n_streams, self.world_size)
        smax = max(sizes)
        m = y_local.shape[1]
        pad = y_local
        if y_local.shape[0] < smax:
            pad = torch.zeros((smax, m), dtype=y_local.dtype, device=y_local.device)
            pad[: y_local.shape[0]] = y_local
        # complex tensors travel as their float32 view (NCCL has no complex type)
        flat = torch.view_as_real(pad.contiguous()).reshape(-1)
        if dst is None:
            out = torch.empty(self.world_size * flat.numel(), dtype=flat.dtype, device=flat.device)
            dist.all_gather_into_tensor(out, flat)
        else:
            bufs = [torch.empty_like(flat) for _ in range(self.world_size)] if self.rank == dst else None
            dist.gather(flat, bufs, dst=dst)
            if self.rank != dst:
                return None
            out = torch.stack(bufs)
        out = torch.view_as_complex(out.reshape(self.world_size, smax, m, 2))
        return torch.cat([out[r, : sizes[r]] for r in range(self.world_size)], dim=0)


class MultiDeviceDDC:
    """Single-process driver: one DigitalDownConverter per device, streams sharded by `shard_range`, host arrays in/out."""

    def __init__(self, decimation_factor: int, sampling_frequency: float, ddc_coeff_filename: str,
                 devices: Sequence[int] | None = None):
        from .ddc import DigitalDownConverter

        if devices is None:
            import torch

            devices = list(range(torch.cuda.device_count()))
        if not devices:
            raise RuntimeError("MultiDeviceDDC needs at least one CUDA device (no CPU fallback)")
        self.devices = list(devices)
        self.workers = [DigitalDownConverter(decimation_factor, sampling_frequency, ddc_coeff_filename, device=d)
                        for d in self.devices]

    def run_batch(self, input_data: np.ndarray, center_freq: float) -> np.ndarray:
        x = np.ascontiguousarray(input_data, dtype=np.float32)
        if x.ndim != 2:
            raise ValueError("run_batch needs a [streams, N] array")
        s, n = x.shape
        m = self.workers[0].out_len(n)
        out = np.empty((s, m), dtype=np.complex64)
        errs: list[BaseException] = []

        def work(i):
            a, b = shard_range(s, len(self.workers), i)
            if a == b:
                return
            try:
                self.workers[i].run_batch(x[a:b], center_freq, out=out[a:b])
            except BaseException as e:  # propagate to the caller thread
                errs.append(e)

        threads = [threading.Thread(target=work, args=(i,)) for i in range(len(self.workers))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errs:
            raise errs[0]
        return out

    def close(self):
        for w in self.workers:
            w.close()
