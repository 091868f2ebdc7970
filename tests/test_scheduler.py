"""CPU tests of the stream scheduler (no CUDA): sharding arithmetic and the torch.distributed gather on the gloo
backend with world_size 2 and 3 (unequal shards), with an injected stand-in for the per-rank compute."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dc_sand_b200.scheduler import ShardedDDC, shard_range, shard_sizes


def test_shard_range_partitions_every_stream_once():
    for n in (0, 1, 7, 64, 128, 129):
        for w in (1, 2, 3, 4, 8):
            seen = []
            for r in range(w):
                a, b = shard_range(n, w, r)
                assert 0 <= a <= b <= n
                seen += list(range(a, b))
            assert seen == list(range(n))
            sizes = shard_sizes(n, w)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
    assert shard_range(128, 8, 3) == (48, 64)
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _fake_ddc(x):
    """Stand-in for the CUDA operator: [s, N] real -> [s, N // 4] complex64, a deterministic function of the input."""
    y = x[:, ::4].to(torch.float32)
    return torch.complex(y, -2.0 * y)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_streams, n, dst, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_streams * n, dtype=torch.float32).reshape(n_streams, n)
        sh = ShardedDDC(n_streams, rank, world, compute=_fake_ddc)
        y_local = sh.run_local(full[sh.start : sh.stop])
        got = sh.gather(y_local, dst=dst)
        ref = _fake_ddc(full)
        if dst is None or rank == dst:
            ok = got is not None and got.shape == ref.shape and torch.equal(got, ref)
        else:
            ok = got is None
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_streams,dst", [(2, 8, None), (2, 5, 0), (3, 7, None)])
def test_sharded_gather_over_gloo(world, n_streams, dst):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, 64, dst, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(world))
    assert res == {r: True for r in range(world)}


def test_run_local_checks_shard_size():
    sh = ShardedDDC(8, 1, 2, compute=_fake_ddc)
    assert list(sh.local_streams) == [4, 5, 6, 7]
    with pytest.raises(ValueError):
        sh.run_local(torch.zeros(3, 16))


def test_device_placement_deals_ranks_over_both_halves():
    """Fewer ranks than devices: ranks alternate between the two host-bridge groups (0 .. n/2 - 1 and n/2 .. n - 1); as many
    ranks as devices: identity; an explicit order overrides; every rank of a job gets a distinct device."""
    from dc_sand_b200.scheduler import device_for_local_rank as dev

    assert [dev(r, 2, 8) for r in range(2)] == [0, 4]
    assert [dev(r, 4, 8) for r in range(4)] == [0, 4, 1, 5]
    assert [dev(r, 8, 8) for r in range(8)] == list(range(8))
    assert [dev(r, 3, 4) for r in range(3)] == [0, 2, 1]
    assert [dev(r, 2, 3) for r in range(2)] == [0, 2]
    assert dev(0, 1, 1) == 0 and dev(0, 1, 8) == 0
    for n_dev in range(1, 9):
        for world in range(1, n_dev + 1):
            got = [dev(r, world, n_dev) for r in range(world)]
            assert len(set(got)) == world and all(0 <= d < n_dev for d in got), (n_dev, world, got)
    assert [dev(r, 4, 8, "4,5,6,7,0,1,2,3") for r in range(4)] == [4, 5, 6, 7]
    for bad in ("0,0,1,2", "0,1", "0,1,2,9"):
        with pytest.raises(ValueError):
            dev(0, 4, 8, bad)
    with pytest.raises(ValueError):
        dev(4, 4, 8)


def test_weighted_shard_sizes():
    """Host-link-aware partition of the end-to-end leg: proportional to the measured rates, every stream assigned once, at
    least one stream per rank, and a symmetric box (rates within the tolerance) keeps the balanced shards."""
    from dc_sand_b200.scheduler import shard_sizes, weighted_shard_range, weighted_shard_sizes

    assert weighted_shard_sizes(128, [25, 25, 25, 25, 47, 47, 47, 47]) == [11, 11, 11, 11, 21, 21, 21, 21]
    assert weighted_shard_sizes(128, [50, 51, 49, 50]) == shard_sizes(128, 4)          # within 15 %: balanced
    assert weighted_shard_sizes(128, [1, 3]) == [32, 96]
    assert weighted_shard_sizes(10, [1, 1000]) == [1, 9]                                # nobody is left without a stream
    assert weighted_shard_sizes(3, [1, 100, 1, 1]) == shard_sizes(3, 4)                 # fewer streams than ranks
    rng = np.random.default_rng(5)
    for _ in range(200):
        world = int(rng.integers(1, 9))
        n = int(rng.integers(world, 300))
        w = rng.uniform(5.0, 60.0, size=world)
        sizes = weighted_shard_sizes(n, w)
        assert sum(sizes) == n and min(sizes) >= 1
        spans = [weighted_shard_range(n, w, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        if max(w) / min(w) > 1.15:
            t = [s / v for s, v in zip(sizes, w)]        # time per rank ~ streams / rate: within one stream of the ideal
            ideal = n / w.sum()
            assert max(t) <= ideal + 1.0 / min(w) + 1e-9
    for bad in ([], [1.0, 0.0], [1.0, float("nan")], [1.0, -2.0]):
        with pytest.raises(ValueError):
            weighted_shard_sizes(8, bad)

