"""Whole-file inference front-end sharded by time: equals the one-shot module result bit for bit, and the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

MEL_KW = dict(sr=16000, win_length=2048, n_mels=229, hop_length=512, fmin=30, fmax=8000, verbose=False)


def test_time_sharded_file_equals_single_pass_and_oracle():
    import reconvat_b200 as R
    from reconvat_b200 import synth, transcribe
    from oracle.frontend import FrontEndOracle
    dev = torch.device("cuda:0")
    L = 700 * 512 + 333                                        # ~22 s, not a multiple of anything
    a16 = np.concatenate([synth.music_int16(L // 2, 31), synth.white_int16(L - L // 2, 32) // 8])
    mel = R.Spectrogram.MelSpectrogram(**MEL_KW).to(dev)
    whole = mel.normalised_log_mel(torch.from_numpy(a16).to(dev)[None, :])          # (1,1,T,229), module path
    T = whole.shape[2]
    one, span = transcribe.whole_file_frontend(mel, torch.from_numpy(a16), 0, 1)
    assert span == (0, T) and torch.equal(one, whole)
    ref = FrontEndOracle().spec_for_model(synth.to_float(a16)[None, :])
    assert np.abs(whole.cpu().numpy() - ref).max() < 1e-4
    for world in (2, 3, 8):
        # emulate the ranks one after the other: pass 1 collects every rank's keys, the reduction is a MAX
        keys = []
        for r in range(world):
            transcribe.whole_file_frontend(mel, torch.from_numpy(a16), r, world,
                                           reduce_keys=lambda k: (keys.append(k.clone()), k)[1])
        wide = torch.stack([k.to(torch.int64) & 0xFFFFFFFF for k in keys]).max(0).values
        glob = torch.where(wide >= 2 ** 31, wide - 2 ** 32, wide).to(torch.int32)
        parts, spans = [], []
        for r in range(world):
            s, sp = transcribe.whole_file_frontend(mel, torch.from_numpy(a16), r, world, reduce_keys=lambda k: glob)
            parts.append(s)
            spans.append(sp)
        assert spans[0][0] == 0 and spans[-1][1] == T and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert torch.equal(torch.cat(parts, dim=2), whole)


def test_one_hour_file_over_8_ranks_is_bit_identical():
    """BASELINE config 5 at full size: 3 600 s = 57.6 M samples = 112 500 frames, time-sharded over 8 ranks."""
    import reconvat_b200 as R
    from reconvat_b200 import synth, transcribe
    dev = torch.device("cuda:0")
    L = 3600 * 16000
    a16 = torch.from_numpy(np.tile(synth.music_int16(16000 * 60, 77), 60)[:L])
    a16[L // 2:] //= 4
    mel = R.Spectrogram.MelSpectrogram(**MEL_KW).to(dev)
    whole, span = transcribe.whole_file_frontend(mel, a16, 0, 1)
    assert span == (0, 112500) and whole.shape == (1, 1, 112500, 229)
    assert float(whole.min()) == 0.0 and float(whole.max()) == 1.0        # exact extrema over the whole file
    keys = []
    for r in range(8):
        transcribe.whole_file_frontend(mel, a16, r, 8, reduce_keys=lambda k: (keys.append(k.clone()), k)[1])
    wide = torch.stack([k.to(torch.int64) & 0xFFFFFFFF for k in keys]).max(0).values
    glob = torch.where(wide >= 2 ** 31, wide - 2 ** 32, wide).to(torch.int32)
    f = 0
    for r in range(8):
        s, (f0, f1) = transcribe.whole_file_frontend(mel, a16, r, 8, reduce_keys=lambda k: glob)
        assert f0 == f and torch.equal(s, whole[:, :, f0:f1])
        f = f1
    assert f == 112500


def test_padded_slice_on_device_equals_the_host_version():
    from reconvat_b200 import transcribe
    dev = torch.device("cuda:0")
    a = torch.randint(-32768, 32767, (5000,), dtype=torch.int16)
    for pinned in (a, a.pin_memory()):
        for s0, s1 in ((0, 5000 + 2 * 64), (0, 700), (10, 900), (64, 3000), (3000, 5000 + 2 * 64), (4000, 5100), (100, 5127)):
            want = transcribe.padded_slice(pinned, s0, s1, 64)
            got = transcribe._padded_slice_on_device(pinned, s0, s1, 64, dev)
            torch.cuda.synchronize()
            assert torch.equal(got.cpu(), want), (s0, s1)
    f = torch.randn(300)
    assert torch.equal(transcribe._padded_slice_on_device(f, 0, 300 + 16, 8, dev).cpu(), transcribe.padded_slice(f, 0, 316, 8))


def test_padded_slice_matches_reflection_pad():
    from reconvat_b200 import transcribe
    a = torch.arange(50, dtype=torch.float32)
    want = torch.nn.functional.pad(a[None, None, :], (8, 8), mode="reflect")[0, 0]
    assert torch.equal(transcribe.padded_slice(a, 0, 66, 8), want)
    assert torch.equal(transcribe.padded_slice(a, 5, 60, 8), want[5:60])
    assert np.array_equal(transcribe.padded_slice(a.numpy(), 5, 60, 8), want[5:60].numpy())
    with pytest.raises(AssertionError):
        transcribe.padded_slice(a[:5], 0, 21, 8)


def test_chunked_batched_transcription_matches_the_whole_file_pass():
    """f3 driver (transcribe_files.py:12-40): the reference's real ``UNet`` (random init, eval) on one file --
    ``UNet.transcribe`` in one batch-1 piece vs ``transcribe_file``: overlapping 640-frame windows through the network
    as a batch, interior frames stitched.  Then the same file over three emulated ranks (contiguous runs of windows,
    min / max keys MAX-reduced): identical to the one-rank result bit for bit.  And the decoded notes agree."""
    import refmodels as RM
    from reconvat_b200 import synth, transcribe
    _, pat = RM.namespaces()
    dev = torch.device("cuda:0")
    frames = 2100
    L = frames * 512
    a16 = np.concatenate([synth.music_int16(L // 2, 41), synth.white_int16(L - L // 2, 42) // 16])
    audio = torch.from_numpy(synth.to_float(a16))
    with RM.deterministic(), torch.no_grad():
        m = RM.build(pat, "unet", dev, 1e-6, 1.3, seed=3).eval()
        whole = m.transcribe({"audio": audio.to(dev)[None, :]})["frame"][0]            # (T, 88), one piece
        pred, (f0, f1) = transcribe.transcribe_file(m, audio, batch=4)
        assert (f0, f1) == (0, frames) and pred["frame"].shape == whole.shape == (frames, 88)
        err = float((pred["frame"] - whole).abs().max())
        assert err < 2e-5, err                                # halo 128 >= the network's receptive field
        narrow = transcribe.transcribe_file(m, audio, halo=16, batch=4)[0]["frame"]
        assert float((narrow - whole).abs().max()) > err      # ... and a halo that is too small shows
        # PCM16 input takes the same path
        pcm, _ = transcribe.transcribe_file(m, torch.from_numpy(a16), batch=4)
        assert float((pcm["frame"] - whole).abs().max()) < 2e-5
        # three ranks, emulated: pass 1 collects the keys, the reduction is a MAX
        world, keys = 3, []
        for r in range(world):
            transcribe.transcribe_file(m, audio, batch=4, rank=r, world_size=world, network=lambda s: s.new_zeros(s.shape[0], s.shape[2], 88),
                                       reduce_keys=lambda k: (keys.append(k.clone()), k)[1])
        wide = torch.stack([k.to(torch.int64) & 0xFFFFFFFF for k in keys]).max(0).values
        glob = torch.where(wide >= 2 ** 31, wide - 2 ** 32, wide).to(torch.int32)
        parts, at = [], 0
        for r in range(world):
            p, (g0, g1) = transcribe.transcribe_file(m, audio, batch=4, rank=r, world_size=world, reduce_keys=lambda k: glob)
            assert g0 == at
            at = g1
            parts.append(p["frame"])
        # (windows are independent, but cuDNN's choice of kernel depends on the batch a window runs in: the front-end
        # pieces are bit-identical -- tested above --, the network's posteriors agree to rounding)
        assert at == frames and float((torch.cat(parts) - pred["frame"]).abs().max()) < 1e-5
    from reconvat_b200 import decoding
    n_whole = decoding.extract_notes_wo_velocity(whole, whole)
    n_chunk = decoding.extract_notes_wo_velocity(pred["onset"], pred["frame"])
    assert len(n_whole[0]) == len(n_chunk[0]) and np.array_equal(n_whole[1], n_chunk[1])
