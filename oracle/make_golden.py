"""Regenerate tests/golden/*.npz by running the UNMODIFIED reference.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container:

    python -m oracle.make_golden

Needs /root/reference; the resulting fixtures are committed so that the GPU
box (which has no reference tree) can check both the oracle and the CUDA path
against the reference's own outputs.  All inputs are framework-RNG-free
(reconvat_b200/synth.py) except ``d``, which is stored in the fixture.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.reference_loader import load_reference          # noqa: E402
from reconvat_b200 import synth                               # noqa: E402
from reconvat_b200.standin import StandInTranscriber          # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
MEL_KW = dict(sr=16000, win_length=2048, n_mels=229, hop_length=512, fmin=30, fmax=8000,
              trainable_mel=False, trainable_STFT=False, verbose=False)   # self_attention_VAT.py:1027-1029


def _frontend(ref, mel, audio):
    """UNet.run_on_batch lines 1112-1121 verbatim, on a float tensor (B, L)."""
    spec = mel(audio.reshape(-1, audio.shape[-1])[:, :-1])
    melpow = spec.clone()
    spec = torch.log(spec + 1e-5)
    logmel = spec.clone()
    spec = ref.utils.Normalization("imagewise").transform(spec)
    spec = spec.transpose(-1, -2).unsqueeze(1)
    return melpow.numpy(), logmel.numpy(), spec.contiguous().numpy()


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    torch.use_deterministic_algorithms(True)
    ref = load_reference()
    S = ref.Spectrogram
    os.makedirs(OUT, exist_ok=True)
    mel = S.MelSpectrogram(**MEL_KW)

    # ---- 1. basis tables (sampled rows + banded mel + checksums)
    rows = np.array([0, 1, 2, 4, 5, 255, 511, 512, 767, 1023, 1024])
    mb = mel.mel_basis.numpy()
    nz_m, nz_k = np.nonzero(mb)
    np.savez_compressed(
        os.path.join(OUT, "basis_16k.npz"), rows=rows,
        wsin_rows=mel.stft.wsin[rows, 0].numpy(), wcos_rows=mel.stft.wcos[rows, 0].numpy(),
        window_mask=mel.stft.window_mask.numpy().reshape(-1),
        wsin_sum64=mel.stft.wsin.double().sum().item(), wcos_sum64=mel.stft.wcos.double().sum().item(),
        wsin_abs64=mel.stft.wsin.double().abs().sum().item(), wcos_abs64=mel.stft.wcos.double().abs().sum().item(),
        mel_nz_m=nz_m.astype(np.int32), mel_nz_k=nz_k.astype(np.int32), mel_nz_v=mb[nz_m, nz_k],
        mel_shape=np.array(mb.shape))

    # ---- 2. short front-end cases (33 frames): white, music, impulse, quiet, all-zeros
    L = 16385
    a16 = np.stack([synth.white_int16(L, 11), synth.music_int16(L, 12), synth.impulse_int16(L, 5000),
                    (synth.music_int16(L, 13) // 64).astype(np.int16), np.zeros(L, np.int16)])
    melpow, logmel, spec = _frontend(ref, mel, torch.from_numpy(synth.to_float(a16)))
    np.savez_compressed(os.path.join(OUT, "frontend_short.npz"), audio_int16=a16, mel_power=melpow,
                        log_mel=logmel, spec=spec)

    # ---- 3. one full 327 680-sample segment per kind (inputs regenerated from synth by seed)
    full = np.stack([synth.white_int16(synth.SEGMENT_SAMPLES, 21), synth.music_int16(synth.SEGMENT_SAMPLES, 22)])
    melpow, logmel, spec = _frontend(ref, mel, torch.from_numpy(synth.to_float(full)))
    np.savez_compressed(os.path.join(OUT, "frontend_full.npz"), seeds=np.array([21, 22]),
                        kinds=np.array(["white", "music"]), log_mel=logmel,
                        spec_min_max=np.stack([logmel.reshape(2, -1).min(1), logmel.reshape(2, -1).max(1)]),
                        spec_stride7=spec.reshape(2, -1)[:, ::7].copy(),
                        audio_crc=np.array([int(full[i].astype(np.int64).sum()) for i in range(2)]))

    # ---- 4. minimum legal length: Spectrogram.py:214-215 admits L-1 == 1024 but torch ReflectionPad1d
    #         needs pad < length, so the shortest input that runs is L-1 == 1025
    a_min = synth.white_int16(1026, 31)[None]
    melpow, logmel, spec = _frontend(ref, mel, torch.from_numpy(synth.to_float(a_min)))
    np.savez_compressed(os.path.join(OUT, "frontend_minlen.npz"), audio_int16=a_min, mel_power=melpow, spec=spec)

    # ---- 5. STFT module, all three output formats, default + small config
    out = {}
    x_s = torch.from_numpy(synth.to_float(np.stack([synth.white_int16(8192, 41), synth.music_int16(8192, 42)])))
    out["audio_int16"] = np.stack([synth.white_int16(8192, 41), synth.music_int16(8192, 42)])
    for tag, kw in (("default", dict(n_fft=2048, hop_length=512, sr=16000)),
                    ("small", dict(n_fft=512, hop_length=128, sr=16000)),
                    ("hop_default", dict(n_fft=1024, sr=22050)),
                    ("nocenter", dict(n_fft=512, hop_length=256, center=False)),
                    ("constpad", dict(n_fft=512, hop_length=128, pad_mode="constant")),
                    ("win_short", dict(n_fft=512, win_length=400, hop_length=160)),
                    ("hamming", dict(n_fft=512, hop_length=128, window="hamming"))):
        st = S.STFT(verbose=False, **kw)
        out[tag + "_complex"] = st(x_s, output_format="Complex").numpy()
        out[tag + "_magnitude"] = st(x_s, output_format="Magnitude").numpy()
        out[tag + "_phase"] = st(x_s, output_format="Phase").numpy()
    np.savez_compressed(os.path.join(OUT, "stft_formats.npz"), **out)

    # ---- 6. MelSpectrogram with other constructor arguments (generic path)
    out = {"audio_int16": out["audio_int16"]}
    for tag, kw in (("librosa_default", dict(verbose=False)),
                    ("htk_128", dict(sr=16000, n_fft=1024, n_mels=128, hop_length=256, htk=True, verbose=False)),
                    ("power1", dict(sr=16000, n_fft=512, n_mels=40, hop_length=128, power=1.0, fmin=20, fmax=7000, verbose=False))):
        out[tag] = S.MelSpectrogram(**kw)(x_s).numpy()
    np.savez_compressed(os.path.join(OUT, "mel_variants.npz"), **out)

    # ---- 7. VAT flavours on a stand-in network, d injected via the seeded global RNG and stored
    B, T, F, P = 2, 24, 229, 24
    xs = synth.uniform01(B * T * F, 51).reshape(B, 1, T, F).astype(np.float32)
    xs.reshape(B, -1)[:, 0] = 0.0          # imagewise normalisation always yields an exact 0 and 1
    xs.reshape(B, -1)[:, 1] = 1.0
    vat_cases = {
        "unet": (ref.self_attention_VAT.UNet_VAT, dict(XI=1e-6, epsilon=2, n_power=1, KL_Div=False), "unet", 4),
        "unet_eps13": (ref.self_attention_VAT.UNet_VAT, dict(XI=1e-6, epsilon=1.3, n_power=1, KL_Div=False), "unet", 4),
        "stepwise_sa": (ref.self_attention_VAT.stepwise_VAT, dict(XI=1e-6, epsilon=2, n_power=1, KL_Div=False), "stepwise", 4),
        "stepwise_vatpy": (ref.VAT.stepwise_VAT, dict(XI=1e-6, epsilon=2, n_power=1), "stepwise", 3),
        "unet_onset": (ref.UNet_onset.UNet_VAT, dict(XI=1e-6, epsilon=2, n_power=1, KL_Div=False), "unet_onset", 4),
        "onf": (ref.onset_frame_VAT.stepwise_VAT, dict(XI=1e-6, epsilon=0.1, n_power=1, KL_Div=False), "onf", 3),
        "unet_kl": (ref.self_attention_VAT.UNet_VAT, dict(XI=1e-6, epsilon=2, n_power=1, KL_Div=True), "unet", 4),
        # the flavours no shipped script selects
        "stepwise_sa_kl": (ref.self_attention_VAT.stepwise_VAT, dict(XI=1e-6, epsilon=2, n_power=1, KL_Div=True), "stepwise", 4),
        "stepwise_sa_binwise": (ref.self_attention_VAT.stepwise_VAT,
                                dict(XI=1e-6, epsilon=2, n_power=1, KL_Div=False, binwise=True), "stepwise", 4),
        "onf_kl": (ref.onset_frame_VAT.stepwise_VAT, dict(XI=1e-6, epsilon=0.1, n_power=1, KL_Div=True), "onf", 3),
        "seg": (ref.Segmentation.Seg_VAT, dict(XI=1e-6, epsilon=2, n_power=1, KL_Div=False), "seg", 4),
        "stack_activation": (ref.onset_frame_VAT.stepwise_VAT_frame_stack,
                             dict(XI=1e-6, epsilon=2, n_power=1, VAT_mode="activation"), "stack", 3),
        "stack_frame": (ref.onset_frame_VAT.stepwise_VAT_frame_stack,
                        dict(XI=1e-6, epsilon=2, n_power=1, VAT_mode="frame"), "stack", 3),
        "stack_all": (ref.onset_frame_VAT.stepwise_VAT_frame_stack,
                      dict(XI=1e-6, epsilon=2, n_power=1, VAT_mode="all"), "stack", 3),
    }
    # With the shipped XI=1e-6 the perturbed and clean posteriors differ at fp32 rounding level, so g
    # (and hence r_adv) is only reproducible on the same device with the same kernels: those cases pin
    # the kernels with g injected.  The "_xi01" twins (XI=0.1) are well conditioned and pin whole modules
    # across devices.
    for tag in list(vat_cases):
        cls, kw, conv, xdim = vat_cases[tag]
        vat_cases[tag + "_xi01"] = (cls, dict(kw, XI=0.1), conv, xdim)
    out = {"x": xs, "P": np.array(P)}
    for tag, (cls, kw, conv, xdim) in vat_cases.items():
        model = StandInTranscriber(conv, n_in=F, n_out=P, seed=3)
        model.captured_grads = []
        x = torch.from_numpy(xs if xdim == 4 else xs[:, 0])
        torch.manual_seed(1234)
        d = torch.randn_like(x)
        torch.manual_seed(1234)
        res = cls(**kw)(model, x)
        vat_loss, r_adv = res[0], res[1]
        out[tag + "_d"] = d.numpy()
        if isinstance(vat_loss, dict):
            out[tag + "_loss"] = np.array([vat_loss["frame"].item(), vat_loss["onset"].item()])
        else:
            out[tag + "_loss"] = np.array([vat_loss.item()])
        out[tag + "_r_adv"] = r_adv.detach().numpy()
        if len(res) > 2:
            out[tag + "_dhat"] = res[2].detach().numpy()
        out[tag + "_g"] = np.stack([g.numpy() for g in model.captured_grads])   # [power steps..., (final pass if backward ran)]
        # gradient of vat_loss w.r.t. the network weights (what training consumes)
        tot = sum(vat_loss.values()) if isinstance(vat_loss, dict) else vat_loss
        model.captured_grads = None
        model.zero_grad()
        tot.backward()
        out[tag + "_wgrad"] = model.frame.weight.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "vat_flavours.npz"), **out)

    # ---- 8. Normalization (utils.py:82-106), both modes, incl. a constant image
    l = torch.from_numpy(logmel[:, :, :40].copy())
    l = torch.cat([l, torch.full_like(l[:1], -3.0)])
    np.savez_compressed(os.path.join(OUT, "normalization.npz"), x=l.numpy(),
                        imagewise=ref.utils.Normalization("imagewise").transform(l).numpy(),
                        framewise=ref.utils.Normalization("framewise").transform(l.clone()).numpy())

    # ---- 9. local-window attention of the U-Net: the reference class, forward and backward.
    # Weights come from the integer hash (no framework RNG), so only inputs' seeds and the outputs are stored.
    from reconvat_b200.standin import _hash_normal
    out = {}
    for tag, (B_, L_, fin, cout, W_, G_, pos) in {"small": (2, 37, 20, 48, 7, 4, True),
                                                  "nopos": (1, 19, 12, 24, 5, 2, False),
                                                  "unet": (1, 70, 229, 916, 31, 4, True)}.items():
        att_mod = ref.self_attention_VAT.MutliHeadAttention1D(fin, cout, W_, position=pos, groups=G_)
        hn = lambda n, seed, shape: torch.tensor(_hash_normal(n, seed).reshape(shape), dtype=torch.float32)
        with torch.no_grad():
            att_mod.W_q.weight.copy_(hn(cout * fin, 301, (cout, fin)) / np.sqrt(fin))
            att_mod.W_k.weight.copy_(hn(cout * fin, 302, (cout, fin)) / np.sqrt(fin))
            att_mod.W_v.weight.copy_(hn(cout * fin, 303, (cout, fin)) / np.sqrt(fin))
            if pos:
                att_mod.rel.copy_(hn(cout * W_, 304, (1, cout, W_)))
        xa = hn(B_ * L_ * fin, 305, (B_, L_, fin)).requires_grad_(True)
        o, a = att_mod(xa)
        go = hn(B_ * L_ * cout, 306, (B_, L_, cout))
        o.backward(go)
        out[tag + "_dims"] = np.array([B_, L_, fin, cout, W_, G_, int(pos)])
        out[tag + "_out"] = o.detach().numpy()
        out[tag + "_att"] = a.detach().numpy()
        out[tag + "_dx"] = xa.grad.numpy()
        if pos:
            out[tag + "_drel"] = att_mod.rel.grad.numpy()
        if tag != "unet":
            for nm in ("W_q", "W_k", "W_v"):
                out[tag + "_d" + nm] = getattr(att_mod, nm).weight.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "attention.npz"), **out)

    # ---- 10. note decoding (model/decoding.py:4-55), both rules, on run-structured random posteriors
    T_, P_ = 300, 88
    u = synth.uniform01(T_ * P_, 401).reshape(T_, P_)
    fr_roll = np.zeros((T_, P_), np.float32)
    on_roll = np.zeros((T_, P_), np.float32)
    starts = np.argwhere(u > 0.985)
    lens = (synth.uniform01(len(starts), 402) * 40).astype(int) + 1
    for (t0, p0), n in zip(starts, lens):
        fr_roll[t0:t0 + n, p0] = 0.9
        on_roll[t0:t0 + 1 + n // 8, p0] = 0.8
    fr_roll[-5:, 3] = 0.9; on_roll[-5, 3] = 0.8                       # a note that runs into the end of the file
    on_roll[10, 7] = 0.8                                               # an onset without a frame (rule1 drops it)
    noise = (synth.uniform01(T_ * P_, 403).reshape(T_, P_) * 0.45).astype(np.float32)
    on_roll = np.maximum(on_roll, noise); fr_roll = np.maximum(fr_roll, noise[::-1].copy())
    out = {"onsets": on_roll, "frames": fr_roll}
    for rule in ("rule1", "rule2"):
        pch, itv = ref.decoding.extract_notes_wo_velocity(torch.from_numpy(on_roll), torch.from_numpy(fr_roll), 0.5, 0.5, rule=rule)
        out[rule + "_pitches"], out[rule + "_intervals"] = pch, itv
        tt, ff = ref.decoding.notes_to_frames(pch, itv, (T_, P_))
        out[rule + "_frame_counts"] = np.array([len(f) for f in ff])
        out[rule + "_frame_bins"] = np.concatenate(ff) if len(ff) else np.array([])
    pch, itv = ref.decoding.extract_notes_wo_velocity(torch.from_numpy(on_roll), torch.from_numpy(fr_roll), 0.95, 0.95)
    out["none_pitches"], out["none_intervals"] = pch, itv              # thresholds nothing reaches: empty arrays
    np.savez_compressed(os.path.join(OUT, "decoding.npz"), **out)

    for f in sorted(os.listdir(OUT)):
        print("%-28s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == "__main__":
    main()
