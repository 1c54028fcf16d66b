"""The reference's CPU path for the whole hot-path step, restated with the same ATen calls.

TEST INFRASTRUCTURE (see oracle/__init__.py): used as the checker in tests and as the timed CPU
baseline by ``bench.py`` (``cpu_baseline`` and ``--impl reference``).  The reference is pure Python, so
there is nothing to compile into ``oracle/_ref``; and /root/reference is absent on the GPU box, so the
baseline that travels is this port (``cpu_baseline.kind == "port"``).  It issues the op sequence of

  front-end   model/Spectrogram.py:208-231, :456-460; model/self_attention_VAT.py:1100-1104;
              model/utils.py:93-100
  VAT         model/self_attention_VAT.py:162-202 (UNet_VAT.forward), through torch autograd

on CPU tensors with all host threads torch is given.
"""
import torch
import torch.nn.functional as F

from .frontend import FrontEndOracle


class CpuHotPath:
    def __init__(self, xi=1e-6, eps=2.0, **frontend_kw):
        fo = FrontEndOracle(**frontend_kw)
        self.wsin = torch.from_numpy(fo.wsin).unsqueeze(1)          # (F, 1, n_fft) like the registered buffers
        self.wcos = torch.from_numpy(fo.wcos).unsqueeze(1)
        self.mel_basis = torch.from_numpy(fo.mel_basis)
        self.n_fft, self.hop = fo.n_fft, fo.hop
        self.xi, self.eps = xi, eps

    def frontend(self, audio):
        """(B, L) float32 -> (B, 1, T, n_mels)."""
        x = audio.reshape(-1, audio.shape[-1])[:, :-1][:, None, :]    # self_attention_VAT.py:1100, broadcast_dim
        x = F.pad(x, (self.n_fft // 2, self.n_fft // 2), mode="reflect")           # Spectrogram.py:216-218
        spec_imag = F.conv1d(x, self.wsin, stride=self.hop)                          # :219
        spec_real = F.conv1d(x, self.wcos, stride=self.hop)                          # :220
        spec = torch.sqrt(spec_real.pow(2) + spec_imag.pow(2)) ** 2.0               # :227,231,458
        spec = torch.matmul(self.mel_basis, spec)                                    # :460
        spec = torch.log(spec + 1e-5)                                                # self_attention_VAT.py:1102
        size = spec.shape                                                            # utils.py:94-100
        x_max = spec.view(size[0], size[1] * size[2]).max(1, keepdim=True)[0].unsqueeze(1)
        x_min = spec.view(size[0], size[1] * size[2]).min(1, keepdim=True)[0].unsqueeze(1)
        spec = (spec - x_min) / (x_max - x_min)
        return spec.transpose(-1, -2).unsqueeze(1)                                   # :1104

    @staticmethod
    def _l2_normalize(d):
        return d / torch.norm(d, dim=-1, keepdim=True)                               # :240-246

    def vat(self, model, x):
        """UNet_VAT.forward, self_attention_VAT.py:162-202.  Returns (vat_loss, r_adv, dhat)."""
        with torch.no_grad():
            y_ref, _ = model.transcriber(x)
        d = torch.randn_like(x, requires_grad=True)
        r = self.xi * self._l2_normalize(d)
        x_adv = (x + r).clamp(0, 1)
        y_pred, _ = model.transcriber(x_adv)
        loss = F.binary_cross_entropy(y_pred, y_ref)
        loss.backward()
        d = d.grad.detach() * 1e10
        model.zero_grad()
        r_adv = self.eps * self._l2_normalize(d)
        assert torch.isnan(r_adv).any() == False, "r_adv has nan"                    # noqa: E712  (:189)
        x_adv = (x + r_adv).clamp(0, 1)
        y_pred, _ = model.transcriber(x_adv)
        vat_loss = F.binary_cross_entropy(y_pred, y_ref)
        return vat_loss, r_adv, self._l2_normalize(d)

    def step(self, model, audio):
        spec = self.frontend(audio)
        vat_loss, r_adv, r_norm = self.vat(model, spec)
        return vat_loss, r_norm.abs().mean()                                         # :1147-1150
