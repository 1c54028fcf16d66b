"""Note decoding (model/decoding.py): oracle pinned to the unmodified reference's outputs, the
device kernel against both -- integer / index work, bit-exact."""
import numpy as np
import pytest
import torch

from oracle import decoding as OD


def _check(g, rule, pitches, intervals, frames_fn):
    assert np.array_equal(pitches, g[rule + "_pitches"]) and pitches.dtype == g[rule + "_pitches"].dtype
    assert np.array_equal(intervals, g[rule + "_intervals"]) and intervals.shape == g[rule + "_intervals"].shape
    t, f = frames_fn(pitches, intervals, (300, 88))
    assert np.array_equal(t, np.arange(300))
    assert np.array_equal(np.array([len(x) for x in f]), g[rule + "_frame_counts"])
    assert np.array_equal(np.concatenate(f), g[rule + "_frame_bins"])


@pytest.mark.parametrize("rule", ["rule1", "rule2"])
def test_oracle_decoding_matches_reference_golden(golden, rule):
    g = golden["decoding"]
    p, i = OD.extract_notes_wo_velocity(torch.from_numpy(g["onsets"]), torch.from_numpy(g["frames"]), 0.5, 0.5, rule=rule)
    _check(g, rule, p, i, OD.notes_to_frames)
    p0, i0 = OD.extract_notes_wo_velocity(torch.from_numpy(g["onsets"]), torch.from_numpy(g["frames"]), 0.95, 0.95)
    assert p0.shape == g["none_pitches"].shape == (0,) and i0.shape == (0,)


def test_notes_to_frames_vectorised_matches_oracle(golden):
    from reconvat_b200 import decoding as D
    g = golden["decoding"]
    for rule in ("rule1", "rule2"):
        t1, f1 = D.notes_to_frames(g[rule + "_pitches"], g[rule + "_intervals"], (300, 88))
        t2, f2 = OD.notes_to_frames(g[rule + "_pitches"], g[rule + "_intervals"], (300, 88))
        assert np.array_equal(t1, t2) and len(f1) == len(f2) == 300
        assert all(np.array_equal(a, b) for a, b in zip(f1, f2))
    t0, f0 = D.notes_to_frames(np.array([]), np.array([]), (7, 88))
    assert len(f0) == 7 and all(len(x) == 0 for x in f0)


@pytest.mark.gpu
@pytest.mark.parametrize("rule", ["rule1", "rule2"])
def test_decoding_kernel_matches_reference_golden(golden, rule):
    from reconvat_b200 import decoding as D
    dev = torch.device("cuda:0")
    g = golden["decoding"]
    on, fr = torch.from_numpy(g["onsets"]).to(dev), torch.from_numpy(g["frames"]).to(dev)
    p, i = D.extract_notes_wo_velocity(on, fr, 0.5, 0.5, rule=rule)
    _check(g, rule, p, i, D.notes_to_frames)
    p0, i0 = D.extract_notes_wo_velocity(on, fr, 0.95, 0.95, rule=rule)
    assert p0.shape == (0,) and i0.shape == (0,) and p0.dtype == np.float64      # np.array([]) like the reference
    with pytest.raises(NameError):
        D.extract_notes_wo_velocity(on, fr, rule="rule3")
    from reconvat_b200 import _lib
    with pytest.raises(_lib.RvbError):
        D.extract_notes_wo_velocity(on.cpu(), fr.cpu())


@pytest.mark.gpu
def test_decoding_one_hour_roll_against_oracle():
    """112 500 frames x 88 pitches (BASELINE config 5): the kernel against the oracle's Python loops on the same rolls."""
    from reconvat_b200 import decoding as D, synth
    dev = torch.device("cuda:0")
    T, P = 112500, 88
    u = synth.uniform01(T * P, 77).reshape(T, P)
    fr = np.zeros((T, P), np.float32); on = np.zeros((T, P), np.float32)
    starts = np.argwhere(u > 0.9995)
    lens = (synth.uniform01(len(starts), 78) * 60).astype(int) + 1
    for (t0, p0), n in zip(starts, lens):
        fr[t0:t0 + n, p0] = 0.9
        on[t0:t0 + 2, p0] = 0.7
    p, i = D.extract_notes_wo_velocity(torch.from_numpy(on).to(dev), torch.from_numpy(fr).to(dev))
    pr, ir = OD.extract_notes_wo_velocity(torch.from_numpy(on), torch.from_numpy(fr))
    assert len(p) > 3000 and np.array_equal(p, pr) and np.array_equal(i, ir)
