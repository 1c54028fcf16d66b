"""Front-end of whole-file inference, sharded by time (BASELINE config 5).

``UNet.transcribe`` (model/self_attention_VAT.py:1293-1305, driven by transcribe_files.py:63-78) pushes ONE file of
any length through ``spectrogram(audio[:, :-1])`` -> ``log(spec + 1e-5)`` -> ``Normalization('imagewise')`` (min / max over
the WHOLE file, :1302) -> ``transpose``.  A one-hour file is 112 500 frames; here every rank takes a contiguous range of
frames, reads the matching slice of the reflect-padded signal plus a right halo of n_fft - hop samples, runs the same
kernels as the training path on it, and the only exchange is one MAX all-reduce of two uint32 keys per file.  The
concatenation of the ranks' outputs is bit-identical to the single-GPU result (``tests/test_gpu_transcribe.py``).
"""
import numpy as np
import torch

from . import parallel


def padded_slice(audio, s0, s1, pad):
    """Samples [s0, s1) of ReflectionPad1d(pad)(audio) (edge sample not repeated, model/Spectrogram.py:216-218) for a
    1-D numpy array / tensor, built from index arithmetic on the host: only the two file ends ever reflect."""
    n = len(audio)
    idx = np.arange(s0, s1, dtype=np.int64) - pad
    idx = np.where(idx < 0, -idx, idx)
    idx = np.where(idx >= n, 2 * (n - 1) - idx, idx)
    if idx.size and (idx.min() < 0 or idx.max() >= n):
        raise AssertionError("Signal length shorter than reflect padding length (n_fft // 2).")
    if isinstance(audio, torch.Tensor):
        return audio[torch.from_numpy(idx)]
    return audio[idx]


def _padded_slice_on_device(audio, s0, s1, pad, dev):
    """``padded_slice`` without touching the samples on the host: the part of [s0, s1) that lies inside the file goes
    host -> device in ONE copy (asynchronous when ``audio`` is pinned), and the at most ``pad`` reflected samples at
    either file end are gathered on the device from what was just copied.  (Index arithmetic on the host over a
    one-hour file -- 57.6 M samples -- cost 0.7 s per pass; the copy is 2 ms.)"""
    n = len(audio)
    if not isinstance(audio, torch.Tensor):
        audio = torch.from_numpy(np.ascontiguousarray(audio))
    lo, hi = max(s0 - pad, 0), min(s1 - pad, n)               # file samples [lo, hi) are the un-reflected part
    if hi <= lo or s0 - pad < -pad or (s1 - pad) - n > pad or n <= pad:
        return padded_slice(audio, s0, s1, pad).to(dev)       # degenerate slices: the general (host) path
    out = torch.empty(s1 - s0, dtype=audio.dtype, device=dev)
    a = lo + pad - s0                                         # where file sample `lo` lands in the slice
    out[a:a + hi - lo].copy_(audio[lo:hi], non_blocking=True)
    if a > 0:                                                 # left end: padded index i < pad <-> file sample pad - i
        i = torch.arange(s0, s0 + a, device=dev)
        src = pad - i                                         # file samples 1 .. pad, all inside [lo, hi) here
        if pad - s0 >= hi:                                   # (largest source index; host arithmetic, no sync)
            return padded_slice(audio, s0, s1, pad).to(dev)
        out[:a] = out[src - lo + a]
    b = a + hi - lo
    if b < s1 - s0:                                           # right end: file sample j >= n <-> 2 (n - 1) - j
        j = torch.arange(s0 + b, s1, device=dev) - pad
        src = 2 * (n - 1) - j
        if 2 * (n - 1) - (s1 - 1 - pad) < lo:
            return padded_slice(audio, s0, s1, pad).to(dev)
        out[b:] = out[src - lo + a]
    return out


def whole_file_frontend(mel, audio, rank=0, world_size=1, group=None, log_offset=1e-5, trim_last=True,
                        channel_dim=True, reduce_keys=None, frames=None):
    """This rank's frames of the normalised log-Mel image of one long file.

    ``mel``: a CUDA ``reconvat_b200.Spectrogram.MelSpectrogram``; ``audio``: the whole file as a 1-D host tensor /
    array (float32 or PCM int16 -- every rank reads only its own slice).  Returns ``(spec, (f0, f1))`` with ``spec`` of
    shape (1, 1, f1 - f0, n_mels) holding frames [f0, f1) of the file, normalised with the whole-file min / max.
    ``reduce_keys`` replaces the all-reduce of the (1, 2) int32 min/max keys (default:
    ``parallel.global_minmax_keys`` over ``group``).  ``frames=(fa, fb)`` overrides the even split: this rank computes
    frames [fa, fb) -- ranges of different ranks may overlap (min / max are idempotent) as long as together they cover
    the file (the chunked network driver below asks for its windows plus their halos)."""
    dev = mel.mel_basis.device
    if trim_last:
        audio = audio[:-1]                                    # model/self_attention_VAT.py:1296
    n_fft, hop, pad = mel.n_fft, mel.stride, mel.n_fft // 2
    if not (mel.center and mel.pad_mode == "reflect"):
        raise NotImplementedError("whole_file_frontend: the reference's transcribe path uses center=True, reflect")
    if len(audio) <= pad:
        raise AssertionError("Signal length shorter than reflect padding length (n_fft // 2).")
    n_frames = (len(audio) + 2 * pad - n_fft) // hop + 1
    if frames is None:
        f0, f1, s0, s1 = parallel.time_shards(n_frames, hop, n_fft, world_size)[rank]
    else:
        f0, f1 = max(0, min(int(frames[0]), n_frames)), max(0, min(int(frames[1]), n_frames))
        s0, s1 = f0 * hop, ((f1 - 1) * hop + n_fft if f1 > f0 else f0 * hop)
    reduce_keys = reduce_keys or (lambda k: parallel.global_minmax_keys(k, group))
    if f1 <= f0:
        # a short file on many ranks: nothing to compute here, but the collective still needs this rank
        # (all-zero keys are the identity of the MAX reduction)
        reduce_keys(torch.zeros((1, 2), dtype=torch.int32, device=dev))
        n_mels = mel.mel_basis.shape[0]
        shape = (1, 1, 0, n_mels) if channel_dim else (1, 0, n_mels)
        return torch.empty(shape, dtype=torch.float32, device=dev), (f0, f0)
    chunk = _padded_slice_on_device(audio, s0, s1, pad, dev)[None, :]
    spec = mel.normalised_log_mel(chunk, trim_last=False, log_offset=log_offset, channel_dim=channel_dim,
                                  prepadded=True,
                                  reduce_minmax=reduce_keys)
    return spec, (f0, f1)


# ---------------------------------------------------------------------------------------------------------------------
# Chunked, batched whole-file inference (transcribe_files.py:12-40, model/self_attention_VAT.py:1293-1314)
#
# The reference pushes a whole file through the network as ONE batch-1 sequence: its local attention unfolds
# (1, T, 916, 31) tensors -- 12.8 GB each for a one-hour file -- and nothing runs in parallel.  The network is local in
# time (3x3 convolutions on four scales, stride-2 down / up sampling, a 31-frame attention window; BatchNorm is an
# affine map in eval mode), so a frame's posterior only depends on a bounded neighbourhood: the file is cut into
# windows of ``segment`` frames that overlap by 2 * ``halo``, the windows run through the network as a batch, and from
# every window only the frames at least ``halo`` away from a cut are kept (the true file ends are real boundaries and
# keep everything).  Window starts are multiples of 16 frames, the network's total down-sampling factor, so the
# pooling grids of a window and of the whole file coincide.  Windows are independent: ranks take contiguous runs of
# them, and the only exchange is the front-end's min / max all-reduce.

def window_plan(n_frames, segment=640, halo=128):
    """[(w0, w1, k0, k1), ...]: the network sees frames [w0, w1) of the file, frames [k0, k1) of its output are kept.
    The kept ranges tile [0, n_frames) exactly."""
    if segment % 16 or halo % 16 or segment <= 2 * halo:
        raise ValueError("segment and halo must be multiples of 16 with segment > 2 * halo (got %d, %d)" % (segment, halo))
    if n_frames <= segment:
        return [(0, n_frames, 0, n_frames)]
    step, plan, w0 = segment - 2 * halo, [], 0
    while True:
        w1 = min(w0 + segment, n_frames)
        last = w1 == n_frames
        plan.append((w0, w1, 0 if w0 == 0 else w0 + halo, n_frames if last else w1 - halo))
        if last:
            return plan
        w0 += step


def transcribe_file(model, audio, mel=None, segment=640, halo=128, batch=16, rank=0, world_size=1, group=None,
                    network=None, log_offset=1e-5, reduce_keys=None):
    """``UNet.transcribe`` (model/self_attention_VAT.py:1293-1314) for one long file, chunked and batched.

    ``model``: the reference's (patched or not) ``UNet`` in eval mode -- ``model.spectrogram`` must be a
    ``reconvat_b200.Spectrogram.MelSpectrogram`` unless ``mel`` is given; ``network(spec) -> posterior (B, T, 88)``
    defaults to ``model.transcriber(spec)[0]`` (the first piano roll, the only output ``transcribe`` returns; the
    reference also runs its reconstruction branch and throws the result away).  ``audio``: the whole file, 1-D host
    tensor / array, float32 or PCM int16.

    Returns ``(predictions, (f0, f1))``: ``predictions['frame']`` / ``['onset']`` hold this rank's frames [f0, f1) of the
    file's posterior, shape (f1 - f0, 88); with ``world_size == 1`` that is the whole file, what the reference returns
    after its ``squeeze_(0)``."""
    mel = mel if mel is not None else model.spectrogram
    network = network or (lambda s: model.transcriber(s)[0])
    n = len(audio) - 1                                        # the reference drops the last sample (:1296)
    n_frames = (n + 2 * (mel.n_fft // 2) - mel.n_fft) // mel.stride + 1
    plan = window_plan(n_frames, segment, halo)
    lo, hi = parallel.segment_shard(len(plan), rank, world_size)
    mine = plan[lo:hi]
    fa, fb = (mine[0][0], mine[-1][1]) if mine else (0, 0)
    spec, (fa, fb) = whole_file_frontend(mel, audio, rank, world_size, group, log_offset, frames=(fa, fb),
                                         reduce_keys=reduce_keys)
    dev = spec.device
    out = []
    with torch.no_grad():
        for i in range(0, len(mine), batch):
            group_ = mine[i:i + batch]
            full = [w for w in group_ if w[1] - w[0] == segment]
            rest = [w for w in group_ if w[1] - w[0] != segment]       # the (shorter) last window of the file
            post = {}
            if full:
                x = torch.stack([spec[0, :, w[0] - fa:w[1] - fa] for w in full])       # (b, 1, segment, n_mels)
                y = network(x)
                for w, yw in zip(full, y):
                    post[w] = yw
            for w in rest:
                post[w] = network(spec[:, :, w[0] - fa:w[1] - fa])[0]
            for w in group_:
                out.append(post[w][w[2] - w[0]:w[3] - w[0]])
    n_out = 88 if not out else out[0].shape[-1]
    roll = torch.cat(out) if out else torch.empty((0, n_out), dtype=torch.float32, device=dev)
    f0, f1 = (mine[0][2], mine[-1][3]) if mine else (0, 0)
    return {"onset": roll, "frame": roll}, (f0, f1)


def gather_frames(local, f0, f1, n_frames, group=None):
    """All ranks' (f1 - f0, P) pieces -> the (n_frames, P) posterior of the file on every rank (one all-reduce of a
    zero-initialised roll: the pieces are disjoint)."""
    import torch.distributed as dist
    full = torch.zeros((n_frames, local.shape[-1]), dtype=local.dtype, device=local.device)
    full[f0:f1] = local
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(full, group=group)
    return full
