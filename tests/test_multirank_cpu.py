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


# ---- the caller's training step without the host in the loop (reconvat_b200.training) ----
class _TinyStepModel(torch.nn.Module):
    """Has the reference's step interface: run_on_batch(batch_l, batch_ul, VAT) -> (predictions, losses, spec)."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.lin = torch.nn.Linear(6, 3)

    def run_on_batch(self, batch_l, batch_ul=None, VAT=False):
        y = torch.sigmoid(self.lin(batch_l["audio"]))
        losses = {"loss/train_frame": torch.nn.functional.binary_cross_entropy(y, batch_l["frame"])}
        if batch_ul is not None:
            yu = torch.sigmoid(self.lin(batch_ul["audio"]))
            losses["loss/train_LDS_ul"] = ((yu - 0.5) ** 2).mean()
        return {"frame": y}, losses, batch_l["audio"]


def _batches(seed, n):
    g = torch.Generator().manual_seed(seed)
    return [{"audio": torch.randn(4, 6, generator=g), "frame": (torch.rand(4, 3, generator=g) > 0.5).float()} for _ in range(n)]


def _train_worker(rank, world, port, out, use_ddp):
    from reconvat_b200 import training
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _TinyStepModel()
        runner = training.ddp(model) if use_ddp else model
        opt = torch.optim.SGD(model.parameters(), lr=0.5)
        sched = torch.optim.lr_scheduler.StepLR(opt, 1000)
        # every rank trains on ITS half of each global batch
        l, ul = _batches(1, 3), _batches(2, 3)
        mine = lambda bs: [{k: v[2 * rank:2 * rank + 2] for k, v in b.items()} for b in bs]      # noqa: E731
        _, losses, _ = training.train_VAT_model(runner, 3, 0, mine(l), mine(ul), opt, sched, 3.0, alpha=1.0, VAT=True)
        out.put((rank, [p.detach().double().flatten().tolist() for p in model.parameters()], sorted(losses)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("use_ddp", [False, True])
def test_training_loop_averages_gradients_over_two_ranks(use_ddp):
    """Two ranks, each on half of every batch, end with IDENTICAL parameters, equal to one process training on the
    whole batches (means over equal shards average exactly); both the flattened all-reduce and the DDP adapter."""
    from reconvat_b200 import training
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q, use_ddp)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict()
    for _ in procs:
        r, params, keys = q.get()
        got[r] = params
        assert keys == ["loss/train_LDS_ul", "loss/train_frame"]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    single = _TinyStepModel()
    opt = torch.optim.SGD(single.parameters(), lr=0.5)
    training.train_VAT_model(single, 3, 0, _batches(1, 3), _batches(2, 3), opt, torch.optim.lr_scheduler.StepLR(opt, 1000),
                             3.0, alpha=1.0, VAT=True)
    for a, b, c in zip(got[0], got[1], single.parameters()):
        assert a == b
        assert np.allclose(a, c.detach().double().flatten().numpy(), atol=1e-6)
