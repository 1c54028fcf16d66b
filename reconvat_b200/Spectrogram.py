"""Drop-in for ``nnAudio.Spectrogram`` (== the reference's vendored model/Spectrogram.py) limited to
the classes on the hot path: ``STFT`` (model/Spectrogram.py:22-316) and ``MelSpectrogram``
(model/Spectrogram.py:319-466).  Same constructor arguments, same registered buffers
(``wsin``, ``wcos``, ``window_mask``, ``mel_basis`` -- checkpoint compatible, transcribe_files.py:71
loads strictly), same output shapes and formats; the arithmetic runs in librvb.so:

    MelSpectrogram:  pad + frame + fold + block-scaled fp16 hi/lo split (float or PCM16 in)  ->  folded 3xFP16
                     contraction on tcgen05 CTA pairs with |.|^2 and the banded Mel projection in its epilogue
                     [-> log -> per-segment min/max -> normalise -> transpose: ``normalised_log_mel``]
    STFT:            the same contraction with a magnitude / complex / phase epilogue (+ a scalar Nyquist-bin kernel)
    other bases:     3xTF32 folded / unfolded contractions, separate banded Mel kernel (see ``_spectrum``)

Scope cuts, all raising instead of silently computing something else: ``trainable*=True`` (no
config of the reference uses it), ``STFT.inverse`` / the other nnAudio transforms,
CPU tensors (there is no fallback path).
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib, basis

__all__ = ["STFT", "MelSpectrogram"]


def _ceil_div(a, b):
    return -(-a // b)


def _fused_fold_enabled():
    """PCM16 input: fold inside the contraction (K0x + K1x, no materialised frame planes) or materialise the folded
    fp16 planes first (K0q + K1q)?  RVB_FUSED_FOLD=1 / 0 selects; the default is the faster of the two on B200 at
    the benchmark shape (profiles/r02_experiments.md)."""
    v = os.environ.get("RVB_FUSED_FOLD")
    return _FUSED_FOLD_DEFAULT if v is None else v not in ("0", "")


_FUSED_FOLD_DEFAULT = False


class STFT(nn.Module):
    """model/Spectrogram.py:22-237.  ``forward(x, output_format=None)``:
    ``Magnitude`` -> (B, F, T); ``Complex`` -> (B, F, T, 2) holding (re, -im); ``Phase`` -> (B, F, T)."""

    def __init__(self, n_fft=2048, win_length=None, freq_bins=None, hop_length=None, window='hann',
                 freq_scale='no', center=True, pad_mode='reflect', iSTFT=False,
                 fmin=50, fmax=6000, sr=22050, trainable=False,
                 output_format="Complex", verbose=True):
        super().__init__()
        if trainable:
            raise NotImplementedError("reconvat_b200.STFT: trainable=True (autograd through the Fourier basis, "
                                      "model/Spectrogram.py:170-174) is outside the accelerated hot path")
        if win_length is None:
            win_length = n_fft
        if hop_length is None:
            hop_length = int(win_length // 4)
        self.output_format = output_format
        self.trainable = trainable
        self.stride = hop_length
        self.center = center
        self.pad_mode = pad_mode
        self.n_fft = n_fft
        self.freq_bins = freq_bins
        self.pad_amount = self.n_fft // 2
        self.window = window
        self.win_length = win_length
        self.iSTFT = iSTFT

        kernel_sin, kernel_cos, self.bins2freq, self.bin_list, window_mask = basis.fourier_basis(
            n_fft, win_length=win_length, freq_bins=freq_bins, window=window, freq_scale=freq_scale,
            fmin=fmin, fmax=fmax, sr=sr)
        kernel_sin = torch.from_numpy(kernel_sin).unsqueeze(1)        # (F, 1, n_fft)
        kernel_cos = torch.from_numpy(kernel_cos).unsqueeze(1)
        if iSTFT:   # kept only so that state_dict keys match; inverse() is not accelerated
            self.register_buffer('kernel_sin_inv', torch.cat((kernel_sin, -kernel_sin[1:-1].flip(0)), 0).unsqueeze(-1))
            self.register_buffer('kernel_cos_inv', torch.cat((kernel_cos, kernel_cos[1:-1].flip(0)), 0).unsqueeze(-1))
        window_mask = torch.from_numpy(window_mask)
        # float32 product, as model/Spectrogram.py:162-164
        self.register_buffer('wsin', kernel_sin * window_mask)
        self.register_buffer('wcos', kernel_cos * window_mask)
        self.register_buffer('window_mask', window_mask.unsqueeze(0).unsqueeze(-1))
        self._tables = None          # device operand planes, rebuilt lazily from wsin/wcos
        self._tables_key = None
        if verbose:
            print("STFT kernels created (reconvat_b200, tcgen05 split-precision contraction)")

    # -- device tables ------------------------------------------------------------------
    def _device_tables(self):
        key = (self.wsin.device, self.wsin._version, self.wcos._version, self.wsin.data_ptr())
        if self._tables is None or self._tables_key != key:
            if not self.wsin.is_cuda:
                raise _lib.RvbError("reconvat_b200.STFT: module is on %s; move it to a CUDA device "
                                    "(there is no CPU path)" % self.wsin.device)
            wcos = self.wcos[:, 0, :].detach().cpu().numpy()
            wsin = self.wsin[:, 0, :].detach().cpu().numpy()
            dev = self.wsin.device
            tb = dict(n_bins=wcos.shape[0], fold=None, fold2=None, fold2x=None, direct=None)
            # RVB_STFT_OPERAND=tf32 keeps the 3xTF32 planes (half the MMA rate; kept for A/B measurements)
            operand = os.environ.get("RVB_STFT_OPERAND", "f16")
            fold = None if os.environ.get("RVB_NO_FOLD") else basis.fold_operand(wcos, wsin, operand=operand)
            if fold is None and operand == "f16" and not os.environ.get("RVB_NO_FOLD"):
                fold = basis.fold_operand(wcos, wsin, operand="tf32")       # n_fft % 128 != 0
            if fold is not None:
                for k in ("basis_hi", "basis_lo", "left_cos", "left_sin"):
                    fold[k] = torch.from_numpy(np.ascontiguousarray(fold[k])).to(dev)
                tb["fold"] = fold
                # the twice-folded operand of the fused Mel path (RVB_NO_FOLD2=1: keep the once-folded contraction)
                if fold["operand"] == "f16" and not os.environ.get("RVB_NO_FOLD2"):
                    f2 = basis.fold2_operand(wcos, wsin)
                    if f2 is not None:
                        for k in ("basis_hi", "basis_lo"):
                            f2[k] = torch.from_numpy(np.ascontiguousarray(f2[k])).to(dev)
                        tb["fold2"] = f2
                        # ... and its twin for the contraction that folds in-kernel (PCM16 input): half the weight in
                        # the centre column, where the converter's e = p[n] + p[N-n] counts the sample twice
                        fx = basis.fold2_operand(wcos, wsin, centre_doubled=True)
                        for k in ("basis_hi", "basis_lo"):
                            fx[k] = torch.from_numpy(np.ascontiguousarray(fx[k])).to(dev)
                        tb["fold2x"] = fx
            else:
                hi, lo, n_gemm, leftover = basis.gemm_operand(wcos, wsin)
                tb["direct"] = dict(basis_hi=torch.from_numpy(hi).to(dev), basis_lo=torch.from_numpy(lo).to(dev),
                                    n_gemm_bins=n_gemm, leftover=leftover)
            self._tables, self._tables_key = tb, key
        return self._tables

    def _apply(self, fn, *args, **kwargs):
        self._tables = None
        return super()._apply(fn, *args, **kwargs)

    # -- geometry -----------------------------------------------------------------------
    def _geometry(self, num_samples, prepadded=False):
        if self.center and not prepadded:
            if self.pad_mode == 'reflect':
                if num_samples < self.pad_amount:
                    raise AssertionError("Signal length shorter than reflect padding length (n_fft // 2).")
                if num_samples == self.pad_amount:
                    # the reference reaches nn.ReflectionPad1d, which refuses pad >= length
                    raise RuntimeError("Padding size should be less than the corresponding input dimension, but got: "
                                       "padding (%d, %d) at dimension 2 of input [1, 1, %d]"
                                       % (self.pad_amount, self.pad_amount, num_samples))
                mode = _lib.PAD_REFLECT
            elif self.pad_mode == 'constant':
                mode = _lib.PAD_CONSTANT
            else:
                raise ValueError("pad_mode must be 'reflect' or 'constant'")
            padded = num_samples + 2 * self.pad_amount
        else:
            mode, padded = _lib.PAD_NONE, num_samples
        if padded < self.n_fft:
            raise RuntimeError("Calculated padded input size per channel: (%d). Kernel size: (%d). "
                               "Kernel size can't be greater than actual input size" % (padded, self.n_fft))
        n_frames = (padded - self.n_fft) // self.stride + 1
        rows = _ceil_div(padded, self.stride)
        return mode, n_frames, rows

    def _check_input(self, x):
        if not x.is_cuda:
            raise _lib.RvbError("reconvat_b200.STFT: input is on %s; there is no CPU path" % x.device)
        if x.requires_grad:
            raise NotImplementedError("reconvat_b200.STFT: gradients w.r.t. the waveform are not provided")
        if x.dtype not in (torch.float32, torch.int16):
            raise _lib.RvbError("reconvat_b200.STFT: expected float32 (or PCM int16) audio, got %s" % x.dtype)
        x2 = x[:, 0, :]
        return x2 if x2.stride(-1) == 1 else x2.contiguous()

    def n_frames(self, num_samples):
        return self._geometry(num_samples)[1]

    def _spectrum(self, x, epilogue, power, make_out, mel_tab=None, prepadded=False, mel_tab2=None):
        """x: (B,1,L) CUDA float32.  Runs pad/frame/split + the tcgen05 contraction with the given epilogue.
        ``make_out(B, n_frames)`` -> (out tensor, n_out_bins).  Returns (out, n_frames).
        ``mel_tab`` (only with the folded fp16 contraction, see :meth:`fused_mel_ok`): fuse the Mel projection;
        make_out then returns the (B, n_mels, T) tensor and n_mels."""
        x2 = self._check_input(x)
        B, L = x2.shape
        mode, n_frames, rows = self._geometry(L, prepadded)      # prepadded: x already carries its padding / halo
        out, n_out_bins = make_out(B, n_frames)
        tb = self._device_tables()
        pcm16 = x2.dtype == torch.int16
        if pcm16 and not (tb["fold"] is not None and tb["fold"]["operand"] == "f16"):
            x2 = x2.to(torch.float32).div_(32768.0)          # model/dataset.py:62; only the fp16 fold reads PCM16 itself
            pcm16 = False
        ld = x2.stride(0) if B > 1 else L
        if tb["fold"] is not None:
            fd = tb["fold"]
            half, M = self.n_fft // 2, B * n_frames
            p0 = torch.empty((M,), dtype=torch.float32, device=x.device) if fd["w0"] != 0.0 else None
            p0_ptr = None if p0 is None else p0.data_ptr()
            if fd["operand"] == "f16":
                if mel_tab2 is not None and pcm16 and tb["fold2x"] is not None and self.stride % 16 == 0 \
                        and 2 * B * (n_frames * self.stride + self.n_fft) < 2 ** 31 \
                        and _fused_fold_enabled():
                    # K0x / K1x: no materialised frame planes.  The padded PCM16 signal is stored once, split by sample
                    # parity (2 bytes per sample); the contraction's converter warps fold / scale / split in-kernel
                    fx = tb["fold2x"]
                    plane_len = _lib.parity_plane_len(L, self.pad_amount, mode, self.n_fft, self.stride, n_frames)
                    sig = torch.empty((2, B, plane_len), dtype=torch.int16, device=x.device)
                    _lib.call("rvb_pad_parity_pcm16", _lib.ptr(x2, torch.int16), ld, B, L, self.pad_amount, mode,
                              sig.data_ptr(), plane_len)
                    _lib.call("rvb_stft_mel_fused_pcm16", sig.data_ptr(), plane_len, B, n_frames, self.n_fft,
                              self.stride, 1.0 / 32768.0, fx["basis_hi"].data_ptr(), fx["basis_lo"].data_ptr(),
                              fx["scale_inv"], mel_tab2.ctypes.data, n_out_bins, _lib.ptr(out))
                    return out, n_frames
                planes = torch.empty((2, 2, M, half), dtype=torch.float16, device=x.device)  # [hi|lo][e|o][frame][c]
                row_inv = torch.empty((M,), dtype=torch.float32, device=x.device)
                if mel_tab2 is not None:
                    # twice-folded contraction: planes with the even-n columns first, four chains over k = 1 .. N/4
                    f2 = tb["fold2"]
                    if pcm16:
                        _lib.call("rvb_fold_split2_f16_pcm16", _lib.ptr(x2, torch.int16), ld, 1.0 / 32768.0, B, L,
                                  self.pad_amount, mode, self.n_fft, self.stride, n_frames, planes[0].data_ptr(),
                                  planes[1].data_ptr(), row_inv.data_ptr())
                    else:
                        _lib.call("rvb_fold_split2_f16", _lib.ptr(x2), ld, B, L, self.pad_amount, mode, self.n_fft,
                                  self.stride, n_frames, planes[0].data_ptr(), planes[1].data_ptr(), row_inv.data_ptr())
                    _lib.call("rvb_stft_mel_folded2_f16", planes[0].data_ptr(), planes[1].data_ptr(), row_inv.data_ptr(),
                              B, n_frames, self.n_fft, f2["basis_hi"].data_ptr(), f2["basis_lo"].data_ptr(),
                              f2["scale_inv"], mel_tab2.ctypes.data, n_out_bins, _lib.ptr(out))
                    return out, n_frames
                if pcm16:
                    _lib.call("rvb_fold_split_f16_pcm16", _lib.ptr(x2, torch.int16), ld, 1.0 / 32768.0, B, L,
                              self.pad_amount, mode, self.n_fft, self.stride, n_frames, planes[0].data_ptr(),
                              planes[1].data_ptr(), row_inv.data_ptr(), p0_ptr)
                else:
                    _lib.call("rvb_fold_split_f16", _lib.ptr(x2), ld, B, L, self.pad_amount, mode, self.n_fft,
                              self.stride, n_frames, planes[0].data_ptr(), planes[1].data_ptr(), row_inv.data_ptr(),
                              p0_ptr)
                if mel_tab is not None:
                    # Mel projection fused into the epilogue: `out` is (B, n_mels, T), the spectrum is never stored
                    _lib.call("rvb_stft_mel_folded_f16", planes[0].data_ptr(), planes[1].data_ptr(),
                              row_inv.data_ptr(), B, n_frames, self.n_fft, fd["basis_hi"].data_ptr(),
                              fd["basis_lo"].data_ptr(), fd["scale_inv"], fd["n_bins_pad"], p0_ptr, fd["w0"], epilogue,
                              float(power), mel_tab.ctypes.data, n_out_bins, _lib.ptr(out))
                    return out, n_frames
                _lib.call("rvb_stft_gemm_folded_f16", planes[0].data_ptr(), planes[1].data_ptr(), row_inv.data_ptr(),
                          B, n_frames, self.n_fft, fd["basis_hi"].data_ptr(), fd["basis_lo"].data_ptr(),
                          fd["scale_inv"], fd["n_bins_pad"], p0_ptr, fd["w0"], epilogue, float(power), _lib.ptr(out),
                          n_out_bins)
                for i, k in enumerate(fd["leftover"]):
                    if k < n_out_bins:
                        _lib.call("rvb_stft_bin_folded_f16", planes[0].data_ptr(), planes[1].data_ptr(),
                                  row_inv.data_ptr(), B, n_frames, self.n_fft, fd["left_cos"][i].data_ptr(),
                                  fd["left_sin"][i].data_ptr(), p0_ptr, fd["w0"], k, epilogue, float(power),
                                  _lib.ptr(out), n_out_bins)
                return out, n_frames
            planes = torch.empty((2, 2, M, half), dtype=torch.float32, device=x.device)    # [hi|lo][e|o][frame][c]
            _lib.call("rvb_fold_split", _lib.ptr(x2), ld, B, L, self.pad_amount, mode, self.n_fft, self.stride,
                      n_frames, planes[0].data_ptr(), planes[1].data_ptr(), None if p0 is None else p0.data_ptr())
            _lib.call("rvb_stft_gemm_folded", planes[0].data_ptr(), planes[1].data_ptr(), B, n_frames, self.n_fft,
                      fd["basis_hi"].data_ptr(), fd["basis_lo"].data_ptr(), fd["n_bins_pad"],
                      None if p0 is None else p0.data_ptr(), fd["w0"], epilogue, float(power), _lib.ptr(out),
                      n_out_bins)
            for i, k in enumerate(fd["leftover"]):
                if k < n_out_bins:
                    _lib.call("rvb_stft_bin_folded", planes[0].data_ptr(), planes[1].data_ptr(), B, n_frames,
                              self.n_fft, fd["left_cos"][i].data_ptr(), fd["left_sin"][i].data_ptr(),
                              None if p0 is None else p0.data_ptr(), fd["w0"], k, epilogue, float(power),
                              _lib.ptr(out), n_out_bins)
            return out, n_frames
        # unfolded contraction (non-symmetric window / non-integer bin scale)
        if self.stride % 32 != 0 or self.n_fft % 32 != 0:
            raise NotImplementedError("reconvat_b200.STFT: for a basis that is not symmetric about n_fft/2, hop_length "
                                      "and n_fft must be multiples of 32 (got hop=%d, n_fft=%d)" % (self.stride, self.n_fft))
        dr = tb["direct"]
        # tail boxes of the last segment run past the planes: TMA zero-fills out-of-bounds rows
        sig = torch.empty((2, B * rows, self.stride), dtype=torch.float32, device=x.device)
        _lib.call("rvb_pad_split", _lib.ptr(x2), ld, B, L, self.pad_amount, mode, sig[0].data_ptr(),
                  sig[1].data_ptr(), rows, self.stride)
        _lib.call("rvb_stft_gemm", sig[0].data_ptr(), sig[1].data_ptr(), B, rows, self.stride, n_frames,
                  dr["basis_hi"].data_ptr(), dr["basis_lo"].data_ptr(), dr["basis_hi"].shape[0], self.n_fft,
                  epilogue, float(power), _lib.ptr(out), n_out_bins)
        for k in dr["leftover"]:
            if k < n_out_bins:
                _lib.call("rvb_stft_bin", sig[0].data_ptr(), sig[1].data_ptr(), B, rows, self.stride, n_frames,
                          self.wcos[k, 0].data_ptr(), self.wsin[k, 0].data_ptr(), self.n_fft, k, epilogue,
                          float(power), _lib.ptr(out), n_out_bins)
        return out, n_frames

    def fused_mel_ok(self):
        """True when the contraction that will run is the folded fp16 one (the only one with the Mel epilogue)."""
        fd = self._device_tables()["fold"]
        return fd is not None and fd["operand"] == "f16" and not os.environ.get("RVB_NO_MEL_FUSION")

    def forward(self, x, output_format=None):
        output_format = output_format or self.output_format
        self.num_samples = x.shape[-1]
        x = basis.broadcast_dim(x)
        F = self.wsin.shape[0]
        dev = x.device
        if output_format == 'Magnitude':
            epi, shape = _lib.EPI_MAGNITUDE, lambda B, T: (B, F, T)
        elif output_format == 'Complex':
            epi, shape = _lib.EPI_COMPLEX, lambda B, T: (B, F, T, 2)
        elif output_format == 'Phase':
            epi, shape = _lib.EPI_PHASE, lambda B, T: (B, F, T)
        else:
            return None          # the reference falls through its if/elif chain the same way
        out, _ = self._spectrum(x, epi, 1.0, lambda B, T: (torch.empty(shape(B, T), dtype=torch.float32, device=dev), F))
        return out

    def inverse(self, *args, **kwargs):
        raise NotImplementedError("reconvat_b200.STFT.inverse: iSTFT is outside the accelerated hot path "
                                  "(SURVEY.md section 2, row 1b)")

    def extra_repr(self):
        return 'n_fft={}, Fourier Kernel size={}, iSTFT={}, trainable={}'.format(
            self.n_fft, (*self.wsin.shape,), self.iSTFT, self.trainable)


class MelSpectrogram(nn.Module):
    """model/Spectrogram.py:319-466.  ``forward(x)`` -> (B, n_mels, T) Mel power spectrogram.

    Extension (not in the reference): :meth:`normalised_log_mel` runs the rest of the front-end of
    ``UNet.run_on_batch`` (model/self_attention_VAT.py:1100-1104) fused on the device.

    ``precision`` (extension; default ``$RVB_PRECISION`` or ``"fast"``) picks the contraction:

    * ``"fast"``: the TWICE-folded contraction -- bins k and N/2-k from the parity-split sums Ce +- Co, a quarter of
      the dense contraction's multiply-adds.  The partial sums carry the energy of BOTH bins, and the tensor core adds
      into its fp32 accumulator (TMEM) with truncation, not rounding: a bin inherits an absolute
      error of ~1e-7 of the amplitude of its mirror bin N/2-k (-140 dB).  On white, music-like and PCM-quantised input
      (the BASELINE signals) log-Mel stays within 4.3e-5 of float64; a band that lies more than ~75 dB below the content
      at its mirror frequency (a full-scale 7 kHz tone over silent low bands) can miss the 1e-4 budget by up to 9x --
      still closer to the truth than the reference's own GPU run with PyTorch's default ``cudnn.allow_tf32=True``
      (1.3e-4 .. 5.6e-4 on the same signals, profiles/r02_precision.md).
    * ``"strict"``: the once-folded contraction (every bin accumulated on its own, twice the multiply-adds, the split
      product's small terms accumulated before the leading one): <= 3.2e-5 on every stress signal, the accuracy class
      of the reference's fp32 path, at about twice the front-end time.
    """

    def __init__(self, sr=22050, n_fft=2048, n_mels=128, hop_length=512,
                 window='hann', center=True, pad_mode='reflect', power=2.0, htk=False,
                 fmin=0.0, fmax=None, norm=1, trainable_mel=False, trainable_STFT=False,
                 verbose=True, precision=None, **kwargs):
        super().__init__()
        precision = precision or os.environ.get("RVB_PRECISION", "fast")
        if precision not in ("fast", "strict"):
            raise ValueError("precision must be 'fast' or 'strict', got %r" % (precision,))
        self.precision = precision
        if trainable_mel or trainable_STFT:
            raise NotImplementedError("reconvat_b200.MelSpectrogram: trainable_mel / trainable_STFT "
                                      "(model/Spectrogram.py:430-433) are outside the accelerated hot path")
        self.stride = hop_length
        self.center = center
        self.pad_mode = pad_mode
        self.n_fft = n_fft
        self.power = power
        self.trainable_mel = trainable_mel
        self.trainable_STFT = trainable_STFT
        self.stft = STFT(n_fft=n_fft, freq_bins=None, hop_length=hop_length, window=window,
                         freq_scale='no', center=center, pad_mode=pad_mode, sr=sr, trainable=trainable_STFT,
                         output_format="Magnitude", verbose=verbose, **kwargs)
        mel_basis = basis.mel_filterbank(sr, n_fft, n_mels, fmin, fmax, htk=htk, norm=norm)
        self.register_buffer('mel_basis', torch.from_numpy(mel_basis))
        self._bands = None
        self._bands_key = None
        self._fused = None
        self._fused_key = None
        self._fused2 = None
        self._fused2_key = None
        if verbose:
            print("Mel filter created (reconvat_b200, banded projection)")

    def _apply(self, fn, *args, **kwargs):
        self._bands = None
        self._fused = None
        self._fused2 = None
        return super()._apply(fn, *args, **kwargs)

    def _band_tables(self):
        key = (self.mel_basis.device, self.mel_basis._version, self.mel_basis.data_ptr())
        if self._bands is None or self._bands_key != key:
            lo, ln, w, k_end = basis.band_rows(self.mel_basis.detach().cpu().numpy())
            dev = self.mel_basis.device
            self._bands = dict(lo=torch.from_numpy(lo).to(dev), len=torch.from_numpy(ln).to(dev),
                               w=torch.from_numpy(w).to(dev), max_len=w.shape[0], k_end=k_end)
            self._bands_key = key
        return self._bands

    def _fused_table(self):
        """Device table of the fused Mel epilogue, or None when this module must take the two-kernel path."""
        key = (self.mel_basis.device, self.mel_basis._version, self.mel_basis.data_ptr())
        if self._fused is None or self._fused_key != key:
            tab = None
            if self.stft.fused_mel_ok():
                fd = self.stft._device_tables()["fold"]
                mb = self.mel_basis.detach().cpu().numpy()
                if not any(np.any(mb[:, k] != 0) for k in fd["leftover"]):    # e.g. the Nyquist bin carries no weight
                    tab = basis.mel_epilogue_table(mb, fd["n_bins_pad"])
            # host-side table: the C ABI reads it at launch time and passes it as a kernel parameter
            self._fused = (np.ascontiguousarray(tab) if tab is not None and tab.shape[0] <= 1024 else None,)
            self._fused_key = key
        return self._fused[0]

    def _fused2_table(self):
        """Host table of the twice-folded contraction's Mel epilogue (power spectrogram only), or None."""
        key = (self.mel_basis.device, self.mel_basis._version, self.mel_basis.data_ptr())
        if self._fused2 is None or self._fused2_key != key:
            tab = None
            if self.precision == "fast" and self.stft.fused_mel_ok() and float(self.power) == 2.0 and \
                    self.stft._device_tables().get("fold2") is not None:
                # 64 rows per epilogue group (32 in the four-chain A/B kernel): bands up to 65 bins wide stay at two
                # partial sums per element
                tab = basis.mel_epilogue_table2(self.mel_basis.detach().cpu().numpy(), self.n_fft,
                                                chunk=32 if os.environ.get("RVB_FOLD2_N64") else 64)
            self._fused2 = (None if tab is None else np.ascontiguousarray(tab),)
            self._fused2_key = key
        return self._fused2[0]

    def _spectrum_epilogue(self):
        if float(self.power) == 2.0:
            return _lib.EPI_POWER
        if float(self.power) == 1.0:
            return _lib.EPI_MAGNITUDE
        return _lib.EPI_POWER_P

    def _mel_fused(self, x, tab, prepadded=False):
        """(B,1,L) -> Mel spectrogram through the contraction with the fused Mel epilogue: (mel, None, T) with mel
        (B, n_mels, T), or -- twice-folded contraction -- (mel_c, mel_s, T), the cos^2 and sin^2 parts whose sum is
        the Mel spectrogram."""
        n_mels, dev = self.mel_basis.shape[0], x.device
        tab2 = self._fused2_table()
        lead = (2,) if tab2 is not None else ()
        out, n_frames = self.stft._spectrum(
            x, self._spectrum_epilogue(), self.power,
            lambda B, T: (torch.empty(lead + (B, n_mels, T), dtype=torch.float32, device=dev), n_mels),
            mel_tab=tab, prepadded=prepadded, mel_tab2=tab2)
        return (out[0], out[1], n_frames) if tab2 is not None else (out, None, n_frames)

    def _power_spectrogram(self, x, prepadded=False):
        """(B,1,L) -> power (B, T, n_pow_bins), time-major, holding (sqrt(re^2+im^2))**power for every bin the
        filterbank reads (model/Spectrogram.py:458)."""
        bands = self._band_tables()
        n_pow_bins = -(-bands["k_end"] // 4) * 4             # bins >= k_end carry zero weight; pad to 16 bytes
        if n_pow_bins > self.mel_basis.shape[1]:
            n_pow_bins = bands["k_end"]
        epi = self._spectrum_epilogue()
        dev = x.device
        power, n_frames = self.stft._spectrum(
            x, epi | _lib.EPI_TIME_MAJOR, self.power,
            lambda B, T: (torch.empty((B, T, n_pow_bins), dtype=torch.float32, device=dev), n_pow_bins),
            prepadded=prepadded)
        return power, n_frames, bands

    def _project(self, power, n_frames, bands, log_offset, layout, out, minmax):
        B, _, n_pow_bins = power.shape
        _lib.call("rvb_mel_project", power.data_ptr(), B, n_frames, n_pow_bins, bands["lo"].data_ptr(),
                  bands["len"].data_ptr(), bands["w"].data_ptr(), bands["max_len"], self.mel_basis.shape[0],
                  float(log_offset), layout, out.data_ptr(), None if minmax is None else minmax.data_ptr())

    def forward(self, x):
        x = basis.broadcast_dim(x)
        tab = self._fused_table()
        if tab is not None or self._fused2_table() is not None:
            mel, mel_b, _ = self._mel_fused(x, tab)
            return mel if mel_b is None else torch.add(mel, mel_b)
        power, n_frames, bands = self._power_spectrogram(x)
        out = torch.empty((power.shape[0], self.mel_basis.shape[0], n_frames), dtype=torch.float32, device=x.device)
        self._project(power, n_frames, bands, -1.0, _lib.LAYOUT_BINS_MAJOR, out, None)
        return out

    def normalised_log_mel(self, audio, trim_last=True, log_offset=1e-5, channel_dim=True, normalise=True,
                           return_minmax=False, prepadded=False, reduce_minmax=None):
        """Fused front-end of ``UNet.run_on_batch`` (model/self_attention_VAT.py:1100-1104, 1112-1121):

            spec = self(audio[:, :-1]); spec = log(spec + 1e-5)
            spec = Normalization('imagewise').transform(spec); spec = spec.transpose(-1,-2).unsqueeze(1)

        Returns a contiguous (B, 1, T, n_mels) tensor ((B, T, n_mels) when ``channel_dim=False``, the
        O&F convention of model/onset_frame_VAT.py:647-651).  ``audio`` is (B, L) / (L) / (B,1,L), float32 or
        the dataset's PCM int16 (then scaled by 1/32768 on the device, model/dataset.py:62).

        Time-sharded inference on one long file (reconvat_b200.transcribe): ``prepadded=True`` says that ``audio`` is a
        slice of the already reflect-padded signal including its halo (frames start at sample 0, no padding is
        added), and ``reduce_minmax`` is called on the (B, 2) int32 min/max keys between the two passes so that the
        caller can all-reduce them over the ranks that hold the other slices (model/self_attention_VAT.py:1302
        normalises over the whole file).
        """
        x = basis.broadcast_dim(audio)
        if trim_last:
            x = x[:, :, :-1]                                  # a view; the kernel takes the row stride
        tab = self._fused_table()
        if tab is not None or self._fused2_table() is not None:
            mel, mel_b, n_frames = self._mel_fused(x, tab, prepadded)
            B, n_mels = mel.shape[0], mel.shape[1]
            out = torch.empty((B, n_frames, n_mels), dtype=torch.float32, device=x.device)
            minmax = None
            fpc = (-(-n_frames // 8) + 3) // 4 * 4                   # frames per CTA of the 8-CTA cluster kernel
            fits = fpc <= 128 and fpc * (n_mels | 1) * 4 <= 200 * 1024 and not os.environ.get("RVB_NO_NORM_FUSION")
            if mel_b is not None and not (normalise and reduce_minmax is None and fits):
                mel, mel_b = torch.add(mel, mel_b), None              # the two-pass kernels take one plane
            if normalise and reduce_minmax is None:
                # one pass: a cluster per segment keeps the log-Mel values in shared memory across the min/max
                minmax = torch.empty((B, 2), dtype=torch.int32, device=x.device)
                _lib.call("rvb_logmel_normalise", mel.data_ptr(), None if mel_b is None else mel_b.data_ptr(), B, n_mels,
                          n_frames, float(log_offset), minmax.data_ptr(), out.data_ptr())
            else:
                if normalise:
                    minmax = torch.empty((B, 2), dtype=torch.int32, device=x.device)
                    _lib.call("rvb_logmel_minmax", mel.data_ptr(), B, n_mels * n_frames, float(log_offset),
                              minmax.data_ptr())
                    minmax = reduce_minmax(minmax)                  # e.g. all-reduce over the ranks of a sharded file
                _lib.call("rvb_logmel_transpose", mel.data_ptr(), B, n_mels, n_frames, float(log_offset),
                          None if minmax is None else minmax.data_ptr(), out.data_ptr())
            out = out.unsqueeze(1) if channel_dim else out
            return (out, minmax) if return_minmax else out
        power, n_frames, bands = self._power_spectrogram(x, prepadded)
        B, n_mels = power.shape[0], self.mel_basis.shape[0]
        out = torch.empty((B, n_frames, n_mels), dtype=torch.float32, device=x.device)
        minmax = torch.empty((B, 2), dtype=torch.int32, device=x.device) if normalise else None
        self._project(power, n_frames, bands, log_offset, _lib.LAYOUT_TIME_MAJOR, out, minmax)
        if normalise and reduce_minmax is not None:
            minmax = reduce_minmax(minmax)
        if normalise:
            _lib.call("rvb_normalise", out.data_ptr(), out.data_ptr(), B, n_frames * n_mels, minmax.data_ptr())
        out = out.unsqueeze(1) if channel_dim else out
        return (out, minmax) if return_minmax else out

    def extra_repr(self):
        return 'Mel filter banks size = {}, trainable_mel={}'.format(
            (*self.mel_basis.shape,), self.trainable_mel, self.trainable_STFT)
