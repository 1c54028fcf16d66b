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


def whole_file_frontend(mel, audio, rank=0, world_size=1, group=None, log_offset=1e-5, trim_last=True,
                        channel_dim=True, reduce_keys=None):
    """This rank's frames of the normalised log-Mel image of one long file.

    ``mel``: a CUDA ``reconvat_b200.Spectrogram.MelSpectrogram``; ``audio``: the whole file as a 1-D host tensor /
    array (float32 or PCM int16 -- every rank reads only its own slice).  Returns ``(spec, (f0, f1))`` with ``spec`` of
    shape (1, 1, f1 - f0, n_mels) holding frames [f0, f1) of the file, normalised with the whole-file min / max.
    ``reduce_keys`` replaces the all-reduce of the (1, 2) int32 min/max keys (default:
    ``parallel.global_minmax_keys`` over ``group``)."""
    dev = mel.mel_basis.device
    if trim_last:
        audio = audio[:-1]                                    # model/self_attention_VAT.py:1296
    n_fft, hop, pad = mel.n_fft, mel.stride, mel.n_fft // 2
    if not (mel.center and mel.pad_mode == "reflect"):
        raise NotImplementedError("whole_file_frontend: the reference's transcribe path uses center=True, reflect")
    if len(audio) <= pad:
        raise AssertionError("Signal length shorter than reflect padding length (n_fft // 2).")
    n_frames = (len(audio) + 2 * pad - n_fft) // hop + 1
    f0, f1, s0, s1 = parallel.time_shards(n_frames, hop, n_fft, world_size)[rank]
    reduce_keys = reduce_keys or (lambda k: parallel.global_minmax_keys(k, group))
    if f1 <= f0:
        # a short file on many ranks: nothing to compute here, but the collective still needs this rank
        # (all-zero keys are the identity of the MAX reduction)
        reduce_keys(torch.zeros((1, 2), dtype=torch.int32, device=dev))
        n_mels = mel.mel_basis.shape[0]
        shape = (1, 1, 0, n_mels) if channel_dim else (1, 0, n_mels)
        return torch.empty(shape, dtype=torch.float32, device=dev), (f0, f0)
    chunk = padded_slice(audio, s0, s1, pad)
    chunk = (chunk if isinstance(chunk, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(chunk)))
    chunk = chunk.to(dev, non_blocking=True)[None, :]
    spec = mel.normalised_log_mel(chunk, trim_last=False, log_offset=log_offset, channel_dim=channel_dim,
                                  prepadded=True,
                                  reduce_minmax=reduce_keys)
    return spec, (f0, f1)
