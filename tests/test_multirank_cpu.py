"""world_size-2 gloo tests of the data-parallel host logic (segment sharding, max-over-ranks timing,
whole-file min/max reduction for time-sharded inference).  The hot path itself has no collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from reconvat_b200 import parallel


def test_segment_shard_partitions_every_batch():
    for n in (0, 1, 7, 8, 32, 33):
        for w in (1, 2, 4, 8):
            spans = [parallel.segment_shard(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.segment_shard(8, 2, 2)


def test_time_shards_cover_a_long_file_with_halo():
    n_frames = 112500                                   # 1 hour at hop 512 (BASELINE config 5)
    sh = parallel.time_shards(n_frames, 512, 2048, 8)
    assert sh[0][0] == 0 and sh[-1][1] == n_frames
    for (f0, f1, s0, s1), nxt in zip(sh, sh[1:] + [None]):
        assert (f1 - f0) % 128 == 0 or nxt is None
        assert s1 - s0 == (f1 - f0 - 1) * 512 + 2048     # right halo of n_fft - hop samples
        if nxt:
            assert nxt[0] == f1 and nxt[2] < s1           # next shard re-reads the halo


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) shard 5 segments over 2 ranks, process the local slice, gather -> every segment exactly once
        seg = torch.arange(5, dtype=torch.float32) * 10
        a, b = parallel.segment_shard(5, rank, world)
        local = seg[a:b] + 1
        got = [torch.zeros(3), torch.zeros(3)]
        pad = torch.zeros(3)
        pad[:b - a] = local
        dist.all_gather(got, pad)
        # (2) timing: the reported time is the max over ranks
        t = parallel.max_over_ranks(1.0 + rank)
        # (3) whole-file min/max across time shards, NaN-propagating
        mn, mx = parallel.global_min_max(torch.tensor([float(rank) - 3.0]), torch.tensor([float(rank) + 5.0]))
        nmn, nmx = parallel.global_min_max(torch.tensor([float("nan") if rank == 1 else 0.0]), torch.tensor([1.0]))
        # (4) the same reduction on the kernels' uint32 keys (time-sharded whole-file inference): one MAX all-reduce
        def key(f):                                           # f2key of rvb_common.cuh
            b = int(np.float32(f).view(np.uint32))
            return (~b & 0xFFFFFFFF) if b & 0x80000000 else (b | 0x80000000)

        def as_i32(u):
            return u - 2 ** 32 if u >= 2 ** 31 else u
        lo, hi = (-7.5, 0.25) if rank == 0 else (-1.0, 3.5)    # rank 0 holds the minimum, rank 1 the maximum
        k = torch.tensor([[as_i32(key(-lo)), as_i32(key(hi))], [as_i32(0xFFFFFFFF if rank == 1 else key(2.0)), as_i32(key(1.0))]],
                         dtype=torch.int32)
        kk = parallel.global_minmax_keys(k)
        want = [as_i32(key(7.5)), as_i32(key(3.5)), as_i32(0xFFFFFFFF), as_i32(key(1.0))]
        if rank == 0:
            out.put((torch.cat([got[0][:3], got[1][:2]]).tolist(), t, mn.item(), mx.item(),
                     bool(torch.isnan(nmn).item()), bool(torch.isnan(nmx).item()), kk.flatten().tolist() == want))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    gathered, t, mn, mx, nan_mn, nan_mx, keys_ok = q.get()
    assert keys_ok
    assert gathered == [1.0, 11.0, 21.0, 31.0, 41.0]
    assert t == 2.0
    assert (mn, mx) == (-3.0, 6.0)
    assert nan_mn and nan_mx
