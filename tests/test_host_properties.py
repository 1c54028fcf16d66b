"""Property tests (hypothesis) of the host-side logic: shard arithmetic, padding slices, operand splits, Mel tables."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from reconvat_b200 import basis, parallel, transcribe
from reconvat_b200 import decoding as D
from oracle import decoding as OD


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 200000), st.sampled_from([128, 160, 256, 512]), st.sampled_from([512, 1024, 2048]),
       st.integers(1, 8), st.sampled_from([1, 32, 128]))
def test_time_shards_partition_frames_and_cover_their_samples(n_frames, hop, n_fft, world, mult):
    sh = parallel.time_shards(n_frames, hop, n_fft, world, frames_multiple=mult)
    assert len(sh) == world and sh[0][0] == 0 and sh[-1][1] == n_frames
    for (f0, f1, s0, s1), nxt in zip(sh, sh[1:] + [None]):
        assert 0 <= f0 <= f1 <= n_frames
        if f1 > f0:
            assert s0 == f0 * hop and s1 == (f1 - 1) * hop + n_fft          # first sample of f0 .. last sample of f1-1
        if nxt is not None:
            assert nxt[0] == f1 and (f1 - f0) % mult == 0 or f1 == n_frames


@settings(max_examples=60, deadline=None)
@given(st.integers(10, 400), st.integers(1, 9), st.data())
def test_padded_slice_is_a_window_of_reflection_pad(n, pad, data):
    a = torch.arange(n, dtype=torch.float32) * 0.5 - 3
    full = torch.nn.functional.pad(a[None, None], (pad, pad), mode="reflect")[0, 0]
    s0 = data.draw(st.integers(0, n + 2 * pad - 1))
    s1 = data.draw(st.integers(s0, n + 2 * pad))
    assert torch.equal(transcribe.padded_slice(a, s0, s1, pad), full[s0:s1])


@settings(max_examples=40, deadline=None)
@given(st.lists(st.floats(-4.0, 4.0, allow_nan=False, width=32), min_size=8, max_size=64), st.integers(-30, 8))
def test_f16_split_carries_22_bits_of_the_block_maximum(vals, exp):
    x = np.asarray(vals, np.float64) * 2.0 ** exp
    if not np.any(x):
        return
    hi, lo, inv = basis.f16_split64(x)
    assert hi.dtype == np.float16 and np.isfinite(hi.astype(np.float64)).all()
    rec = (hi.astype(np.float64) + lo.astype(np.float64)) * inv
    assert np.abs(rec - x).max() <= 2.0 ** -21 * np.abs(x).max()
    assert 2.0 ** 14 <= np.abs(hi.astype(np.float64)).max() <= 2.0 ** 15


@settings(max_examples=25, deadline=None)
@given(st.sampled_from([(16000, 2048), (22050, 2048), (16000, 1024), (44100, 4096)]), st.integers(20, 260),
       st.floats(0.0, 100.0), st.booleans())
def test_mel_epilogue_table_reconstructs_every_bank_it_accepts(cfg, n_mels, fmin, htk):
    sr, n_fft = cfg
    mb = basis.mel_filterbank(sr, n_fft, n_mels, fmin, None, htk=htk)
    n_pad = -(-(mb.shape[1] - 1) // 128) * 128
    tab = basis.mel_epilogue_table(mb, n_pad)
    if tab is None:
        return                                                # bank not representable: the module takes the other path
    band0 = tab[:, 2].view(np.int32)
    assert np.all(np.diff(band0) >= 0)
    dense = np.zeros((n_mels + 1, n_pad), np.float32)
    k = np.arange(n_pad)
    np.add.at(dense, (band0, k), tab[:, 0])
    np.add.at(dense, (band0 + 1, k), tab[:, 1])
    assert np.array_equal(dense[:n_mels, :mb.shape[1] - 1], mb[:, :mb.shape[1] - 1][:, :n_pad])
    for m in range(n_mels):                                   # what makes the RED.ADD epilogue order-independent
        nz = np.flatnonzero(mb[m])
        assert len(nz) == 0 or nz[-1] // 32 - nz[0] // 32 <= 1


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 60), st.integers(1, 12), st.integers(0, 2 ** 31 - 1))
def test_notes_to_frames_vectorised_equals_the_loop(T, P, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(0, 12))
    on = rng.integers(0, T, n)
    off = np.minimum(on + rng.integers(1, 10, n), T)
    pit = rng.integers(0, P, n)
    iv = np.stack([on, off], 1) if n else np.array([])
    t1, f1 = D.notes_to_frames(pit if n else np.array([]), iv, (T, P))
    t2, f2 = OD.notes_to_frames(pit if n else np.array([]), iv, (T, P))
    assert np.array_equal(t1, t2) and len(f1) == len(f2)
    assert all(np.array_equal(a, b) for a, b in zip(f1, f2))
