"""Restatement of the nnAudio==0.2.0 helpers the reference front-end imports.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference vendors ``model/Spectrogram.py`` (a copy of nnAudio 0.2.0's
Spectrogram.py) but still pulls three functions from the pip package, which is
absent from /root/reference and from this image (requirements.txt:8,
``nnAudio == 0.2.0``; no lock file):

* ``broadcast_dim``           -- call sites model/Spectrogram.py:208, :456
* ``create_fourier_kernels``  -- call site  model/Spectrogram.py:133-141
* ``mel``                     -- call site  model/Spectrogram.py:421
                                 (nnAudio.librosa_functions.mel == librosa 0.7
                                 ``filters.mel``, Slaney scale, norm=1)

Their published algorithms are restated here in numpy, in float64 with the
same final casts to float32 as the originals, so that the reference's own
``STFT.__init__`` / ``MelSpectrogram.__init__`` produce the buffers they would
with the real package.
"""
import numpy as np
import torch
from scipy.signal import get_window


def broadcast_dim(x):
    """(L) / (B, L) / (B, 1, L) -> (B, 1, L); anything else is a ValueError
    (nnAudio.utils.broadcast_dim; used at model/Spectrogram.py:208,456)."""
    if x.dim() == 2:
        x = x[:, None, :]
    elif x.dim() == 1:
        x = x[None, None, :]
    elif x.dim() == 3:
        pass
    else:
        raise ValueError("Only support input with shape = (batch, len) or shape = (len)")
    return x


def pad_center(data, size):
    """librosa.util.pad_center for 1-D data (zero pad both sides)."""
    n = data.shape[-1]
    lpad = int((size - n) // 2)
    if lpad < 0:
        raise ValueError("Target size ({:d}) must be at least input size ({:d})".format(size, n))
    return np.pad(data, (lpad, int(size - n - lpad)), mode="constant")


def create_fourier_kernels(n_fft, win_length=None, freq_bins=None, fmin=50, fmax=6000, sr=44100,
                           freq_scale="linear", window="hann", verbose=True):
    """nnAudio.utils.create_fourier_kernels (0.2.0).

    Returns (wsin, wcos, bins2freq, binslist, window_mask); wsin/wcos are
    float32 ``(freq_bins, 1, n_fft)`` *un-windowed* sin/cos tables evaluated in
    float64, window_mask is the float32 periodic window centre-padded to n_fft.
    The reference multiplies the window in afterwards, in float32
    (model/Spectrogram.py:162-164).
    """
    if freq_bins is None:
        freq_bins = n_fft // 2 + 1
    if win_length is None:
        win_length = n_fft
    s = np.arange(0, n_fft, 1.0)
    wsin = np.empty((freq_bins, 1, n_fft))
    wcos = np.empty((freq_bins, 1, n_fft))
    bins2freq, binslist = [], []
    window_mask = get_window(window, int(win_length), fftbins=True)
    window_mask = pad_center(window_mask, n_fft)
    if freq_scale == "linear":
        start_bin = fmin * n_fft / sr
        scaling_ind = (fmax - fmin) * (n_fft / sr) / freq_bins
        for k in range(freq_bins):
            bins2freq.append((k * scaling_ind + start_bin) * sr / n_fft)
            binslist.append((k * scaling_ind + start_bin))
            wsin[k, 0, :] = np.sin(2 * np.pi * (k * scaling_ind + start_bin) * s / n_fft)
            wcos[k, 0, :] = np.cos(2 * np.pi * (k * scaling_ind + start_bin) * s / n_fft)
    elif freq_scale == "log":
        start_bin = fmin * n_fft / sr
        scaling_ind = np.log(fmax / fmin) / freq_bins
        for k in range(freq_bins):
            bins2freq.append(np.exp(k * scaling_ind) * start_bin * sr / n_fft)
            binslist.append((np.exp(k * scaling_ind) * start_bin))
            wsin[k, 0, :] = np.sin(2 * np.pi * (np.exp(k * scaling_ind) * start_bin) * s / n_fft)
            wcos[k, 0, :] = np.cos(2 * np.pi * (np.exp(k * scaling_ind) * start_bin) * s / n_fft)
    elif freq_scale == "no":
        for k in range(freq_bins):
            bins2freq.append(k * sr / n_fft)
            binslist.append(k)
            wsin[k, 0, :] = np.sin(2 * np.pi * k * s / n_fft)
            wcos[k, 0, :] = np.cos(2 * np.pi * k * s / n_fft)
    else:
        raise ValueError("Please select the correct frequency scale, 'linear' or 'log'")
    return (wsin.astype(np.float32), wcos.astype(np.float32), bins2freq, binslist,
            window_mask.astype(np.float32))


# ---- librosa 0.7 filters.mel, as copied into nnAudio.librosa_functions ----

def hz_to_mel(frequencies, htk=False):
    frequencies = np.asanyarray(frequencies)
    if htk:
        return 2595.0 * np.log10(1.0 + frequencies / 700.0)
    f_min, f_sp = 0.0, 200.0 / 3
    mels = (frequencies - f_min) / f_sp
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if frequencies.ndim:
        log_t = frequencies >= min_log_hz
        mels[log_t] = min_log_mel + np.log(frequencies[log_t] / min_log_hz) / logstep
    elif frequencies >= min_log_hz:
        mels = min_log_mel + np.log(frequencies / min_log_hz) / logstep
    return mels


def mel_to_hz(mels, htk=False):
    mels = np.asanyarray(mels)
    if htk:
        return 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
    f_min, f_sp = 0.0, 200.0 / 3
    freqs = f_min + f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if mels.ndim:
        log_t = mels >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (mels[log_t] - min_log_mel))
    elif mels >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (mels - min_log_mel))
    return freqs


def mel_frequencies(n_mels=128, fmin=0.0, fmax=11025.0, htk=False):
    min_mel = hz_to_mel(fmin, htk=htk)
    max_mel = hz_to_mel(fmax, htk=htk)
    mels = np.linspace(min_mel, max_mel, n_mels)
    return mel_to_hz(mels, htk=htk)


def mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm=1, dtype=np.float32):
    """librosa 0.7 ``filters.mel``: float32 ``(n_mels, 1 + n_fft//2)`` triangles,
    area-normalised when ``norm == 1`` (the reference always passes norm=1,
    model/Spectrogram.py:398,421)."""
    if fmax is None:
        fmax = float(sr) / 2
    if norm is not None and norm != 1 and norm != np.inf:
        raise ValueError("Unsupported norm: {}".format(repr(norm)))
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=dtype)
    fftfreqs = np.linspace(0, float(sr) / 2, int(1 + n_fft // 2), endpoint=True)
    mel_f = mel_frequencies(n_mels + 2, fmin=fmin, fmax=fmax, htk=htk)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    if norm == 1:
        enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
        weights *= enorm[:, np.newaxis]
    return weights
