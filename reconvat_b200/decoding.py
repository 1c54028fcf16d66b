"""Drop-in for ``model/decoding.py`` of the reference: note decoding after ``UNet.transcribe`` (transcribe_files.py:12-40).

``extract_notes_wo_velocity`` walks every detected onset with a Python ``while`` loop and two ``.item()`` calls per
frame; here the per-pitch scan is one kernel (``rvb_note_offsets``) and the note list is ``torch.nonzero`` -- same
notes, same order (frame-major, then pitch), same numpy return types, bit-exact.  ``notes_to_frames`` is host logic
in the reference as well and is vectorised with numpy.
"""
import numpy as np
import torch

from . import _lib


def extract_notes_wo_velocity(onsets, frames, onset_threshold=0.5, frame_threshold=0.5, rule='rule1'):
    """onsets, frames: CUDA float tensors [frames, bins].  Returns (pitches, intervals) as numpy arrays exactly like
    model/decoding.py:4-55: ``pitches`` (N,) bin indices, ``intervals`` (N, 2) rows (onset_index, offset_index)."""
    if rule not in ('rule1', 'rule2'):
        raise NameError('Please enter the correct rule name')
    for t in (onsets, frames):
        if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2:
            raise _lib.RvbError("reconvat_b200.decoding needs CUDA float32 [frames, bins] tensors; there is no CPU path")
    if onsets.shape != frames.shape:
        raise ValueError("onsets %s and frames %s differ in shape" % (tuple(onsets.shape), tuple(frames.shape)))
    onsets, frames = onsets.contiguous(), frames.contiguous()
    T, P = onsets.shape
    start = torch.empty((T, P), dtype=torch.uint8, device=onsets.device)
    offset = torch.empty((T, P), dtype=torch.int32, device=onsets.device)
    _lib.call("rvb_note_offsets", onsets.data_ptr(), frames.data_ptr(), T, P, float(onset_threshold),
              float(frame_threshold), int(rule == 'rule1'), start.data_ptr(), offset.data_ptr())
    idx = torch.nonzero(start, as_tuple=False)                 # frame-major order, as the reference iterates
    if idx.shape[0] == 0:
        return np.array([]), np.array([])                      # what np.array([]) of the empty lists gives
    off = offset[idx[:, 0], idx[:, 1]].to(torch.int64)
    pitches = idx[:, 1].cpu().numpy()
    intervals = torch.stack((idx[:, 0], off), dim=1).cpu().numpy()
    return pitches, intervals


def notes_to_frames(pitches, intervals, shape):
    """model/decoding.py:111-131: (time, freqs) with ``freqs[t]`` the active bins of frame t."""
    roll = np.zeros(tuple(shape), dtype=np.int32)
    if len(pitches):
        intervals = np.asarray(intervals).reshape(-1, 2)
        delta = np.zeros((shape[0] + 1, shape[1]), dtype=np.int32)
        np.add.at(delta, (intervals[:, 0], np.asarray(pitches)), 1)
        np.add.at(delta, (intervals[:, 1], np.asarray(pitches)), -1)
        roll = np.cumsum(delta, axis=0)[:-1] > 0
    time = np.arange(shape[0])
    t_idx, p_idx = np.nonzero(roll)
    cuts = np.searchsorted(t_idx, np.arange(1, shape[0]))
    freqs = np.split(p_idx, cuts)
    return time, freqs
