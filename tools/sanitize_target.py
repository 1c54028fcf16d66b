#!/usr/bin/env python
"""Small-shape pass over every hot-path kernel, meant to run UNDER compute-sanitizer (tools/sanitize.sh):
front-end on float and PCM16 input (fold/split or the fused contraction, tcgen05 contraction with the Mel epilogue,
cluster normalise, two-pass normalise; the strict route, the STFT module, the 3xTF32 GEMM), the VAT kernels through the module (eager and stats flavours), the divergence
kernels, a CUDA-graph capture + replay of the whole step.  ``__graft_entry__.smoke()`` runs first (it checks the
results against the CPU checker), so a sanitizer run is also a correctness run."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import reconvat_b200 as R                                   # noqa: E402
from reconvat_b200 import synth                             # noqa: E402
from reconvat_b200.pipeline import HotPathStep              # noqa: E402
from reconvat_b200.standin import InjectedTranscriber, StandInTranscriber   # noqa: E402


def main():
    import __graft_entry__ as g
    g.smoke()
    dev = torch.device("cuda:0")
    frames = int(os.environ.get("RVB_SANITIZE_FRAMES", "160"))
    B, L = 3, frames * 512
    pcm = np.stack([synth.white_int16(L, 3), synth.music_int16(L, 4), synth.white_int16(L, 5)])
    mel = R.Spectrogram.MelSpectrogram(sr=16000, win_length=2048, n_mels=229, hop_length=512, fmin=30, fmax=8000,
                                       verbose=False).to(dev)
    a_f = torch.from_numpy(synth.to_float(pcm)).to(dev)
    a_i = torch.from_numpy(pcm).to(dev)
    s_f = mel.normalised_log_mel(a_f)
    s_i = mel.normalised_log_mel(a_i)
    assert torch.equal(s_f, s_i), "PCM16 and float input differ"
    os.environ["RVB_FUSED_FOLD"] = "1"                      # K0x + K1x: the fold done inside the contraction
    s_x = mel.normalised_log_mel(a_i)
    os.environ.pop("RVB_FUSED_FOLD")
    assert float((s_x - s_i).abs().max()) < 1e-5, "fused-fold path differs"
    for pairs in ("2", "4"):                                # K1qm: frame tiles multicast across a cluster
        os.environ["RVB_FOLD2_MC"] = pairs
        s_m = mel.normalised_log_mel(a_i)
        os.environ.pop("RVB_FOLD2_MC")
        assert torch.equal(s_m, s_i), "multicast contraction differs"
    assert bool(torch.isfinite(s_f).all()) and float(s_f.min()) == 0.0 and float(s_f.max()) == 1.0
    # module surface (forward -> log -> Normalization) = the two-pass kernels
    spec = torch.log(mel(a_f[:, :-1]) + 1e-5)
    spec = R.utils.Normalization("imagewise").transform(spec).transpose(-1, -2).unsqueeze(1)
    assert float((spec - s_f).abs().max()) < 1e-5
    # precision="strict": the once-folded CTA-pair contraction (two passes per chain) with the Mel epilogue, the same
    # kernel with the plain epilogue (STFT module), and the raw 3xTF32 GEMM with and without a split contraction
    mel_s = R.Spectrogram.MelSpectrogram(sr=16000, win_length=2048, n_mels=229, hop_length=512, fmin=30, fmax=8000,
                                         verbose=False, precision="strict").to(dev)
    assert float((mel_s.normalised_log_mel(a_i) - s_i).abs().max()) < 1e-3
    stft = R.Spectrogram.STFT(n_fft=2048, hop_length=512, sr=16000, output_format="Complex", verbose=False).to(dev)
    assert bool(torch.isfinite(stft(a_i[:, :8192])).all())
    from reconvat_b200 import linear
    xw = torch.randn(700, 229, device=dev, requires_grad=True)
    ws = [torch.randn(916, 229, device=dev, requires_grad=True) for _ in range(2)]
    sum(y.square().sum() for y in linear.projections(xw, ws)).backward()
    assert bool(torch.isfinite(xw.grad).all())
    # the caller's BatchNorm2d: train forward / backward (vector and scalar paths), eval forward
    from reconvat_b200 import batchnorm
    for shape in ((2, 16, 64, 36), (3, 5, 17, 13)):
        bn = batchnorm.BatchNorm2d(shape[1]).to(dev)
        xb = torch.randn(shape, device=dev, requires_grad=True)
        bn(xb).square().sum().backward()
        assert bool(torch.isfinite(xb.grad).all()) and bool(torch.isfinite(bn.eval()(xb)).all())
    os.environ["RVB_BN_NHWC"] = "1"
    bn = batchnorm.BatchNorm2d(16).to(dev)                   # channels_last kernels (ticketed reduction)
    xb = torch.randn(2, 16, 64, 36, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    for _ in range(2):
        bn(xb).square().sum().backward()
    assert bool(torch.isfinite(xb.grad).all()) and bool(torch.isfinite(bn.eval()(xb)).all())
    # VAT flavours
    for conv, cls, kw in (("unet", "UNet_VAT", dict(KL_Div=False)), ("unet_onset", "UNet_VAT_onset", dict(KL_Div=False)),
                          ("stepwise", "stepwise_VAT", dict(KL_Div=True)), ("stepwise", "stepwise_VAT", dict(KL_Div=False, binwise=True))):
        m = StandInTranscriber(conv, seed=1).to(dev)
        vat = getattr(R.VAT, cls)(XI=0.1, epsilon=2.0, n_power=1, **kw)
        out = vat(m, spec if conv != "onf" else spec.squeeze(1))
        loss = out[0]
        loss = sum(loss.values()) if isinstance(loss, dict) else loss
        loss.backward()
    # the captured step, replayed on two streams
    step = HotPathStep(InjectedTranscriber(B, frames=frames, seed=3).to(dev), dev)
    bufs = [a_i.clone(), a_i.flip(0).contiguous()]
    step.capture(bufs)
    step.replay_many([0, 1, 0, 1], streams=2)
    torch.cuda.synchronize()
    step.check()
    print("sanitize_target ok: launches %d" % R._lib.launch_count())


if __name__ == "__main__":
    main()
