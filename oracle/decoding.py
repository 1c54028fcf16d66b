"""Oracle restatement of model/decoding.py:4-55 (extract_notes_wo_velocity) and :111-131 (notes_to_frames).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Plain Python loops over CPU tensors, the reference's control flow."""
import numpy as np
import torch


def extract_notes_wo_velocity(onsets, frames, onset_threshold=0.5, frame_threshold=0.5, rule='rule1'):
    onsets = (onsets > onset_threshold).cpu().to(torch.uint8)
    frames = (frames > frame_threshold).cpu().to(torch.uint8)
    onset_diff = torch.cat([onsets[:1, :], onsets[1:, :] - onsets[:-1, :]], dim=0) == 1        # :24
    if rule == 'rule1':
        onset_diff = onset_diff & (frames == 1)                                                # :30
    elif rule != 'rule2':
        raise NameError('Please enter the correct rule name')
    on, fr = onsets.numpy(), frames.numpy()
    pitches, intervals = [], []
    for frame, pitch in torch.nonzero(onset_diff, as_tuple=False).tolist():                    # :37
        offset = frame
        while on[offset, pitch] or fr[offset, pitch]:                                          # :45
            offset += 1
            if offset == on.shape[0]:
                break
        if offset > frame:
            pitches.append(pitch)
            intervals.append([frame, offset])
    return np.array(pitches), np.array(intervals)


def notes_to_frames(pitches, intervals, shape):
    roll = np.zeros(tuple(shape))
    for pitch, (onset, offset) in zip(pitches, intervals):
        roll[onset:offset, pitch] = 1
    time = np.arange(roll.shape[0])
    freqs = [roll[t, :].nonzero()[0] for t in time]
    return time, freqs
