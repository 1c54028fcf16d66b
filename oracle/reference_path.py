"""The hot-path step on the reference's OWN classes (unmodified code from the reference tree / its oracle/_ref snapshot).

TEST INFRASTRUCTURE (see oracle/__init__.py): the timed baselines of ``bench.py`` -- ``--impl reference`` and
``cpu_baseline`` on the host cores (``kind: "reference"``), ``gpu_eager_baseline`` on the same B200 in eager PyTorch
(cuDNN conv1d STFT, cuBLAS Mel matmul, ~45 ATen kernels per VAT call: SURVEY.md 2.1 calls this "the bar").

It only strings the reference's public modules together the way ``UNet.run_on_batch`` does
(model/self_attention_VAT.py:1098-1106): ``Spectrogram.MelSpectrogram`` -> ``torch.log(spec + 1e-5)`` ->
``Normalization('imagewise').transform`` -> transposed view -> ``UNet_VAT.forward(model, spec)``.
"""
import torch

from . import reference_loader


def available():
    return reference_loader.available()


class ReferenceHotPath:
    def __init__(self, device="cpu", xi=1e-6, eps=2.0):
        ns = reference_loader.load_reference()
        c = ns.constants
        self.spectrogram = ns.Spectrogram.MelSpectrogram(                     # model/self_attention_VAT.py:1027-1029
            sr=c.SAMPLE_RATE, win_length=c.WINDOW_LENGTH, n_mels=c.N_BINS, hop_length=c.HOP_LENGTH, fmin=c.MEL_FMIN,
            fmax=c.MEL_FMAX, trainable_mel=False, trainable_STFT=False, verbose=False).to(device)
        self.normalize = ns.utils.Normalization("imagewise")                  # :1042
        self.vat_loss = ns.self_attention_VAT.UNet_VAT(xi, eps, 1, False)     # :1044
        self.source = reference_loader.REFERENCE_ROOT

    def frontend(self, audio):
        spec = self.spectrogram(audio.reshape(-1, audio.shape[-1])[:, :-1])   # :1100
        spec = torch.log(spec + 1e-5)                                         # :1102
        spec = self.normalize.transform(spec)                                 # :1103
        return spec.transpose(-1, -2).unsqueeze(1)                            # :1104

    def step(self, model, audio):
        """One Mel+VAT step; returns (vat_loss, mean |d_hat|) like the logged losses (:1147-1150)."""
        spec = self.frontend(audio)
        lds, _, r_norm = self.vat_loss(model, spec)                           # :1106
        return lds, r_norm.abs().mean()
