"""Data-parallel plumbing for the hot path: one process per GPU, segments sharded, no collective on the path.

The only cross-rank operations are (a) the max-over-ranks of a device-timed interval (benchmarking) and
(b) the whole-file min/max of log-Mel for time-sharded inference on one long file
(model/self_attention_VAT.py:1302 normalises over the whole file), a 2-float all-reduce.  Both work on any
``torch.distributed`` backend (NCCL on the GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def bind_to_gpu_numa(device_index):
    """Pin the calling process to the CPU cores NVML reports as local to GPU ``device_index`` (same NUMA node / PCIe
    root), BEFORE it allocates pinned host buffers: first touch then places them in local memory, and eight ranks
    feeding eight GPUs stop sharing one socket's memory controllers.  Returns the number of cores, or 0 when NVML
    is unavailable (nothing is changed then)."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, n_words)
        cpus = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def segment_shard(n_segments, rank, world_size):
    """Contiguous, balanced slice [start, stop) of the batch dimension owned by ``rank``."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d not in [0, %d)" % (rank, world_size))
    base, extra = divmod(n_segments, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def time_shards(n_frames, hop, n_fft, world_size, frames_multiple=128):
    """Shard one long file by time: rank r owns frames [f0, f1) and needs padded-signal samples
    [f0*hop, (f1-1)*hop + n_fft) -- i.e. a right halo of n_fft - hop samples.  Frame counts are multiples of
    ``frames_multiple`` (the GEMM's M tile) except on the last rank.  Returns [(f0, f1, s0, s1), ...]."""
    per = -(-n_frames // world_size)
    per = -(-per // frames_multiple) * frames_multiple
    out = []
    for r in range(world_size):
        f0, f1 = min(r * per, n_frames), min((r + 1) * per, n_frames)
        out.append((f0, f1, f0 * hop, (f1 - 1) * hop + n_fft if f1 > f0 else f0 * hop))
    return out


def max_over_ranks(value, device=None):
    """MAX all-reduce of a python float (device-timed milliseconds)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def global_min_max(local_min, local_max):
    """Whole-file (min, max) from per-rank tensors: all-reduce MIN and MAX (NaN-propagating like torch.min/max:
    a NaN on any rank makes both results NaN)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_min, local_max
    nan = torch.isnan(local_min) | torch.isnan(local_max)
    flag = nan.to(local_min.dtype)
    mn = torch.where(nan, torch.full_like(local_min, float("inf")), local_min)
    mx = torch.where(nan, torch.full_like(local_max, float("-inf")), local_max)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    bad = flag > 0
    nanv = torch.full_like(mn, float("nan"))
    return torch.where(bad, nanv, mn), torch.where(bad, nanv, mx)


def global_minmax_keys(keys, group=None):
    """All-reduce the (B, 2) int32 min/max keys that ``rvb_logmel_minmax`` / ``rvb_mel_project`` produce.  The keys
    are order-preserving uint32 images of (-min, max) with NaN mapped to 0xffffffff, so ONE MAX all-reduce yields the
    whole-file extrema and propagates NaN like torch.min/max -- bit-exact, no float round trip."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return keys
    wide = keys.to(torch.int64) & 0xFFFFFFFF                 # uint32 value; NCCL / gloo have no uint32 MAX
    dist.all_reduce(wide, op=dist.ReduceOp.MAX, group=group)
    return torch.where(wide >= 2 ** 31, wide - 2 ** 32, wide).to(torch.int32)
