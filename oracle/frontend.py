"""Oracle restatement of the Mel front-end (CPU, numpy/torch, fp32 and fp64).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the
reference lines it follows; all paths are relative to /root/reference.

Numerical spec: for a segment ``a`` of 327 680 samples
  s = a[:-1]                                   self_attention_VAT.py:1100
  p = reflect_pad(s, n_fft//2)                 Spectrogram.py:209-218
  re[k,t] = sum_n p[hop*t+n] * wcos[k,n]       Spectrogram.py:219-220
  im[k,t] = sum_n p[hop*t+n] * wsin[k,n]
  P = sqrt(re^2+im^2) ** power                 Spectrogram.py:227,231,458
  M = mel_basis @ P                            Spectrogram.py:460
  l = log(M + 1e-5)                            self_attention_VAT.py:1102
  x = (l - min l) / (max l - min l)  per segment   utils.py:93-100
  out[b,0,t,m] = x[b,m,t]                      self_attention_VAT.py:1104
"""
import numpy as np

from . import nnaudio_restate as R


class FrontEndOracle:
    """Holds the basis tables exactly as the reference's modules build them
    (STFT.__init__ Spectrogram.py:133-178, MelSpectrogram.__init__ :411-435)."""

    def __init__(self, sr=16000, n_fft=2048, n_mels=229, hop_length=512, fmin=30.0, fmax=8000.0,
                 win_length=None, window="hann", power=2.0, htk=False, norm=1):
        win_length = win_length or n_fft
        ksin, kcos, _, _, wmask = R.create_fourier_kernels(
            n_fft, win_length=win_length, freq_bins=None, window=window, freq_scale="no",
            sr=sr, verbose=False)
        # float32 * float32, as torch does at Spectrogram.py:162-164
        self.wsin = (ksin[:, 0, :] * wmask[None, :]).astype(np.float32)   # (F, n_fft)
        self.wcos = (kcos[:, 0, :] * wmask[None, :]).astype(np.float32)
        self.window_mask = wmask
        self.mel_basis = R.mel(sr, n_fft, n_mels, fmin, fmax, htk=htk, norm=norm)  # (n_mels, F) f32
        self.n_fft, self.hop, self.power = n_fft, hop_length, float(power)
        self.n_mels = n_mels

    # -- Spectrogram.py:209-218 (nn.ReflectionPad1d: edge sample not repeated)
    def reflect_pad(self, s):
        pad = self.n_fft // 2
        if s.shape[-1] < pad:
            raise AssertionError("Signal length shorter than reflect padding length (n_fft // 2).")
        return np.pad(s, [(0, 0)] * (s.ndim - 1) + [(pad, pad)], mode="reflect")

    def frames(self, p):
        """(B, Lp) -> strided view (B, T, n_fft); T = (Lp - n_fft)//hop + 1 (conv1d stride)."""
        B, Lp = p.shape
        T = (Lp - self.n_fft) // self.hop + 1
        st = p.strides
        return np.lib.stride_tricks.as_strided(p, (B, T, self.n_fft), (st[0], st[1] * self.hop, st[1]),
                                               writeable=False)

    # -- Spectrogram.py:219-220
    def stft(self, s, dtype=np.float32):
        """s: (B, L) already trimmed by the caller. Returns re, im: (B, F, T)."""
        s = np.ascontiguousarray(s, dtype=dtype)
        fr = self.frames(self.reflect_pad(s))
        re = np.einsum("btn,kn->bkt", fr, self.wcos.astype(dtype), optimize=True)
        im = np.einsum("btn,kn->bkt", fr, self.wsin.astype(dtype), optimize=True)
        return re, im

    # -- Spectrogram.py:226-231 then :458
    def power_spec(self, re, im):
        mag = np.sqrt(re * re + im * im)
        if self.power == 2.0:
            return mag * mag          # torch pow(x, 2.0) is x*x
        if self.power == 1.0:
            return mag
        return mag ** np.asarray(self.power, dtype=mag.dtype)

    # -- Spectrogram.py:460
    def mel_power(self, s, dtype=np.float32):
        re, im = self.stft(s, dtype)
        P = self.power_spec(re, im)
        return np.einsum("mk,bkt->bmt", self.mel_basis.astype(dtype), P, optimize=True)

    # -- self_attention_VAT.py:1102
    @staticmethod
    def log_compress(M):
        return np.log(M + np.asarray(1e-5, dtype=M.dtype))

    # -- utils.py:93-100 ('imagewise')
    @staticmethod
    def normalise_imagewise(l):
        B = l.shape[0]
        mx = l.reshape(B, -1).max(1)[:, None, None]
        mn = l.reshape(B, -1).min(1)[:, None, None]
        with np.errstate(invalid="ignore", divide="ignore"):
            return (l - mn) / (mx - mn)

    # -- utils.py:85-92 ('framewise'; not selected by any shipped script)
    @staticmethod
    def normalise_framewise(l):
        mx = l.max(1, keepdims=True)
        mn = l.min(1, keepdims=True)
        with np.errstate(invalid="ignore", divide="ignore"):
            out = (l - mn) / (mx - mn)
        out[np.isnan(out)] = 0
        return out

    def log_mel(self, s, dtype=np.float32):
        return self.log_compress(self.mel_power(s, dtype))

    def spec_for_model(self, audio, dtype=np.float32, channel_dim=True):
        """audio: (B, L) untrimmed, as the dataset yields it.  Follows
        UNet.run_on_batch self_attention_VAT.py:1112-1121 -> (B,1,T,n_mels)
        (or (B,T,n_mels) for the O&F model, onset_frame_VAT.py:647-651)."""
        l = self.log_mel(audio[:, :-1], dtype)
        x = self.normalise_imagewise(l)
        x = np.ascontiguousarray(np.swapaxes(x, -1, -2))
        return x[:, None] if channel_dim else x
