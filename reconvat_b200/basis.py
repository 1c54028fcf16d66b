"""Host-side builders for the tables the kernels consume.

* windowed Fourier basis  == what nnAudio 0.2.0 ``create_fourier_kernels`` + the window multiply at
  model/Spectrogram.py:133-164 produce (float64 evaluation, float32 storage, float32 window product);
* Slaney / HTK triangular Mel filterbank == ``nnAudio.librosa_functions.mel`` as called at
  model/Spectrogram.py:421, and its banded form (every FFT bin feeds at most two adjacent bands);
* tf32 hi/lo operand planes in the row order the tcgen05 contraction expects.

Pure numpy; the results are registered as module buffers (same names/shapes as the reference, so
checkpoints interchange) plus non-persistent device tables.
"""
import numpy as np
from scipy.signal import get_window

GEMM_TILE_BINS = 128      # one 256-row operand tile = 128 cos rows + 128 sin rows of the same bins


def broadcast_dim(x):
    """(L) / (B,L) / (B,1,L) -> (B,1,L); ValueError otherwise (nnAudio.utils.broadcast_dim)."""
    if x.dim() == 2:
        return x[:, None, :]
    if x.dim() == 1:
        return x[None, None, :]
    if x.dim() == 3:
        return x
    raise ValueError("Only support input with shape = (batch, len) or shape = (len)")


def fourier_basis(n_fft, win_length=None, freq_bins=None, window="hann", freq_scale="no", fmin=50, fmax=6000,
                  sr=22050):
    """Returns (kernel_sin, kernel_cos, bins2freq, binslist, window_mask): float32 un-windowed tables
    (F, n_fft) and the float32 window centre-padded to n_fft."""
    if freq_bins is None:
        freq_bins = n_fft // 2 + 1
    if win_length is None:
        win_length = n_fft
    s = np.arange(0, n_fft, 1.0)
    k = np.arange(freq_bins, dtype=np.float64)
    if freq_scale == "no":
        bins = k
    elif freq_scale == "linear":
        start_bin = fmin * n_fft / sr
        scaling_ind = (fmax - fmin) * (n_fft / sr) / freq_bins
        bins = k * scaling_ind + start_bin
    elif freq_scale == "log":
        start_bin = fmin * n_fft / sr
        scaling_ind = np.log(fmax / fmin) / freq_bins
        bins = np.exp(k * scaling_ind) * start_bin
    else:
        raise ValueError("Please select the correct frequency scale, 'linear' or 'log'")
    bins2freq = bins * sr / n_fft
    arg = (2 * np.pi * bins)[:, None] * s[None, :] / n_fft           # ((2*pi*k)*s)/n_fft in float64
    kernel_sin = np.sin(arg).astype(np.float32)
    kernel_cos = np.cos(arg).astype(np.float32)
    w = get_window(window, int(win_length), fftbins=True)
    lpad = (n_fft - len(w)) // 2
    if lpad < 0:
        raise ValueError("Target size ({:d}) must be at least input size ({:d})".format(n_fft, len(w)))
    window_mask = np.pad(w, (lpad, n_fft - len(w) - lpad)).astype(np.float32)
    return kernel_sin, kernel_cos, list(bins2freq), list(bins), window_mask


# ---------------------------------------------------------------- Mel
def _hz_to_mel(f, htk):
    f = np.asarray(f, dtype=np.float64)
    if htk:
        return 2595.0 * np.log10(1.0 + f / 700.0)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    lin = f / f_sp
    logp = min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep
    return np.where(f >= min_log_hz, logp, lin)


def _mel_to_hz(m, htk):
    m = np.asarray(m, dtype=np.float64)
    if htk:
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm=1):
    """float32 (n_mels, 1 + n_fft//2) triangular filters, area-normalised for norm == 1."""
    if fmax is None:
        fmax = float(sr) / 2
    if norm is not None and norm != 1 and norm != np.inf:
        raise ValueError("Unsupported norm: {}".format(repr(norm)))
    n_mels = int(n_mels)
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0, float(sr) / 2, n_bins, endpoint=True)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(fmin, htk), _hz_to_mel(fmax, htk), n_mels + 2), htk)
    fdiff = np.diff(edges)
    ramps = edges[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0, np.minimum(lower, upper)).astype(np.float32)
    if norm == 1:
        enorm = 2.0 / (edges[2:n_mels + 2] - edges[:n_mels])
        weights = (weights.astype(np.float64) * enorm[:, None]).astype(np.float32)
    return weights


def banded_filterbank(mel_basis):
    """Dense (n_mels, F) -> (band0 int32[F], w0 f32[F], w1 f32[F], k_begin, k_end).

    Bin k contributes w0[k] to band band0[k] and w1[k] to band band0[k]+1.  Raises ValueError when
    the matrix is not a pairwise-overlapping filterbank (a column with non-adjacent or more than two
    non-zeros, or band indices that decrease with k) -- e.g. a trained dense ``mel_basis``.
    """
    mb = np.asarray(mel_basis, dtype=np.float32)
    n_mels, F = mb.shape
    band0 = np.zeros(F, np.int32)
    w0 = np.zeros(F, np.float32)
    w1 = np.zeros(F, np.float32)
    nzcols = np.flatnonzero((mb != 0).any(0))
    if len(nzcols) == 0:
        raise ValueError("mel_basis is all zero")
    k_begin, k_end = int(nzcols[0]), int(nzcols[-1]) + 1
    prev = 0
    for k in range(k_begin, k_end):
        rows = np.flatnonzero(mb[:, k])
        if len(rows) == 0:
            band0[k] = prev
        elif len(rows) == 1:
            r = int(rows[0])
            # keep band0 non-decreasing: a lone weight may sit in either slot of the (band0, band0+1) pair
            if r == prev + 1:
                band0[k], w1[k] = prev, mb[r, k]
            else:
                band0[k], w0[k] = r, mb[r, k]
        elif len(rows) == 2 and rows[1] == rows[0] + 1:
            band0[k], w0[k], w1[k] = int(rows[0]), mb[rows[0], k], mb[rows[1], k]
        else:
            raise ValueError("mel_basis column %d has non-zeros in rows %s: not a banded triangular "
                             "filterbank; the fused Mel kernel cannot represent it" % (k, rows.tolist()))
        if band0[k] < prev:
            raise ValueError("mel_basis band order decreases at bin %d" % k)
        prev = int(band0[k])
    return band0, w0, w1, k_begin, k_end


EPILOGUE_CHUNK_BINS = 32      # bins folded by one epilogue warp group of the contraction (rvb_stft_gemm.cu)


def mel_epilogue_table(mel_basis, n_bins_pad, tile=EPILOGUE_CHUNK_BINS):
    """Table for the Mel projection fused into the contraction's epilogue: float32 [n_bins_pad, 4] rows
    (w0, w1, band0 as int32 bits, 0) from :func:`banded_filterbank`, band0 kept non-decreasing over the padding.
    Returns None when the fused epilogue cannot represent the bank bit-reproducibly: a bin feeding more than two
    (or non-adjacent) bands, weight on a bin the tiled contraction does not produce (>= n_bins_pad), or a band
    straddling more than two 32-bin epilogue chunks (its partial sums would then be added in a run-dependent order)."""
    mb = np.asarray(mel_basis, dtype=np.float32)
    try:
        band0, w0, w1, k_begin, k_end = banded_filterbank(mb)
    except ValueError:
        return None
    if k_end > n_bins_pad:
        return None
    for m in range(mb.shape[0]):
        nz = np.flatnonzero(mb[m])
        if len(nz) and nz[-1] // tile - nz[0] // tile > 1:
            return None
    tab = np.zeros((n_bins_pad, 4), np.float32)
    b = np.zeros(n_bins_pad, np.int32)
    n = min(len(band0), n_bins_pad)
    b[:n] = band0[:n]
    b[:k_begin] = band0[k_begin]                    # leading zero-weight bins: stay on the first band
    b[k_end:] = band0[k_end - 1]                    # trailing ones: stay on the last
    tab[:n, 0] = w0[:n]
    tab[:n, 1] = w1[:n]
    tab[:, 2] = b.view(np.float32)
    assert np.all(np.diff(b) >= 0)
    return tab


def band_rows(mel_basis, max_len=256):
    """Dense (n_mels, F) -> (band_lo int32[n_mels], band_len int32[n_mels], band_w f32[L, n_mels], k_end).

    Row m reads bins [band_lo[m], band_lo[m] + band_len[m]) with weights band_w[:band_len[m], m] (interior zeros
    are kept as zero weights; the table is stored support-position-major so that a warp of bands reads it
    coalesced).  Raises ValueError when a row's support is wider than ``max_len`` bins, e.g. a trained dense
    ``mel_basis`` -- that is not a banded filterbank and the fused Mel kernel would crawl on it."""
    mb = np.asarray(mel_basis, dtype=np.float32)
    n_mels, F = mb.shape
    lo = np.zeros(n_mels, np.int32)
    ln = np.zeros(n_mels, np.int32)
    for m in range(n_mels):
        nz = np.flatnonzero(mb[m])
        if len(nz):
            lo[m], ln[m] = nz[0], nz[-1] - nz[0] + 1
    L = int(ln.max())
    if L == 0:
        raise ValueError("mel_basis is all zero")
    if L > max_len:
        raise ValueError("mel_basis row support of %d bins exceeds %d: not a banded filterbank; the fused Mel "
                         "kernel cannot represent it" % (L, max_len))
    w = np.zeros((L, n_mels), np.float32)
    for m in range(n_mels):
        w[:ln[m], m] = mb[m, lo[m]:lo[m] + ln[m]]
    return lo, ln, w, int((lo + ln).max())


# ---------------------------------------------------------------- tf32 operand planes
def tf32_round(x):
    """Round-to-nearest (ties away) fp32 -> tf32, identical to PTX cvt.rna.tf32.f32 for finite values."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32_split(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    hi = tf32_round(x)
    lo = tf32_round(x - hi)
    return hi, lo


def gemm_operand(wcos, wsin):
    """(F, n_fft) float32 windowed bases -> (basis_hi, basis_lo, n_gemm_bins, leftover_bins).

    Rows are grouped in 256-row tiles: [128 cos rows | 128 sin rows] of the same 128 bins, so that one
    accumulator tile holds re and im of the same bins.  When F % 128 == 1 (the usual n_fft/2+1) the
    last bin is left to the scalar single-bin kernel instead of costing a whole extra tile.
    """
    F, n_fft = wcos.shape
    n_gemm = F - 1 if (F % GEMM_TILE_BINS == 1 and F > 1) else F
    leftover = list(range(n_gemm, F))
    n_tiles = (n_gemm + GEMM_TILE_BINS - 1) // GEMM_TILE_BINS
    mat = np.zeros((n_tiles * 2 * GEMM_TILE_BINS, n_fft), np.float32)
    for j in range(n_tiles):
        lo_bin, hi_bin = j * GEMM_TILE_BINS, min((j + 1) * GEMM_TILE_BINS, n_gemm)
        mat[256 * j:256 * j + (hi_bin - lo_bin)] = wcos[lo_bin:hi_bin]
        mat[256 * j + 128:256 * j + 128 + (hi_bin - lo_bin)] = wsin[lo_bin:hi_bin]
    hi, lo = tf32_split(mat)
    return hi, lo, n_gemm, leftover


def tf32_split64(x64):
    """float64 values -> (hi, lo) float32 tf32 planes with hi + lo == x to ~2^-22 relative."""
    x64 = np.asarray(x64, dtype=np.float64)
    hi = tf32_round(x64.astype(np.float32))
    lo = tf32_round((x64 - hi.astype(np.float64)).astype(np.float32))
    return hi, lo


def f16_split64(x64):
    """float64 values -> (hi, lo, scale_inv): IEEE binary16 planes of x * 2^s with hi + lo == x * 2^s to 2^-22
    relative (2^-25 absolute where lo is subnormal), s chosen so that max|x| 2^s lies in [2^14, 2^15)."""
    x64 = np.asarray(x64, dtype=np.float64)
    mx = float(np.abs(x64).max())
    s = 14 - int(np.floor(np.log2(mx))) if mx > 0 else 0
    xs = np.ldexp(x64, s)
    hi = xs.astype(np.float16)
    lo = (xs - hi.astype(np.float64)).astype(np.float16)
    return hi, lo, float(np.ldexp(1.0, -s))


def fold_operand(wcos, wsin, tol=2.5e-7, operand="tf32"):
    """Folded operand for windows symmetric about n_fft/2, or None when the basis is not symmetric
    (short / non-periodic windows, non-integer 'linear' / 'log' bin scales).

    With wcos[k][N-n] == wcos[k][n] and wsin[k][N-n] == -wsin[k][n] (to fp32 rounding),
        re[k] = w[0] p[0] + sum_{c=0}^{N/2-1} Bc[k][c] e[c],  e[c] = p[c+1] + p[N-c-1]  (e[N/2-1] = p[N/2])
        im[k] =             sum_{c=0}^{N/2-1} Bs[k][c] o[c],  o[c] = p[c+1] - p[N-c-1]  (o[N/2-1] = 0)
    Bc / Bs average the two mirror entries in float64 (they differ by at most an fp32 ulp).
    Returns dict(basis_hi, basis_lo [2*n_bins_pad, N/2]; n_bins_pad; n_gemm_bins; leftover; w0;
    left_cos, left_sin: folded fp32 rows of the leftover bins; scale_inv).  ``operand`` selects the split:
    "tf32" -> float32 planes (scale_inv 1), "f16" -> block-scaled float16 planes (see :func:`f16_split64`).
    """
    F, N = wcos.shape
    if N % (128 if operand == "f16" else 64) != 0:
        return None
    half = N // 2
    c64, s64 = wcos.astype(np.float64), wsin.astype(np.float64)
    fwd = slice(1, half)                       # n = 1 .. N/2-1
    mirror = slice(N - 1, half, -1)            # N-n = N-1 .. N/2+1
    scale = max(float(np.abs(c64).max()), 1e-30)
    if np.abs(c64[:, fwd] - c64[:, mirror]).max() > tol * scale or np.abs(s64[:, fwd] + s64[:, mirror]).max() > tol * scale:
        return None
    if np.abs(s64[:, 0]).max() > tol * scale or np.abs(s64[:, half]).max() > tol * scale:
        return None
    w0 = float(wcos[0, 0])
    if np.abs(c64[:, 0] - w0).max() > tol * scale:     # the n = 0 term must be the same for every bin
        return None
    bc = np.empty((F, half), np.float64)
    bs = np.zeros((F, half), np.float64)
    bc[:, :half - 1] = 0.5 * (c64[:, fwd] + c64[:, mirror])
    bc[:, half - 1] = c64[:, half]
    bs[:, :half - 1] = 0.5 * (s64[:, fwd] - s64[:, mirror])
    n_gemm = F - 1 if (F % GEMM_TILE_BINS == 1 and F > 1) else F
    leftover = list(range(n_gemm, F))
    n_bins_pad = -(-n_gemm // GEMM_TILE_BINS) * GEMM_TILE_BINS
    mat = np.zeros((2 * n_bins_pad, half), np.float64)
    mat[:n_gemm] = bc[:n_gemm]
    mat[n_bins_pad:n_bins_pad + n_gemm] = bs[:n_gemm]
    if operand == "f16":
        hi, lo, scale_inv = f16_split64(mat)
    elif operand == "tf32":
        (hi, lo), scale_inv = tf32_split64(mat), 1.0
    else:
        raise ValueError("operand must be 'f16' or 'tf32'")
    return dict(basis_hi=hi, basis_lo=lo, n_bins_pad=n_bins_pad, n_gemm_bins=n_gemm, leftover=leftover, w0=w0,
                left_cos=bc[n_gemm:].astype(np.float32), left_sin=bs[n_gemm:].astype(np.float32),
                scale_inv=scale_inv, operand=operand)


FOLD2_TILE_K = 128            # k-values per tile of the twice-folded contraction (rvb_stft_gemm.cu)


def fold2_operand(wcos, wsin, tol=2.5e-7, centre_doubled=False):
    """Operand of the TWICE-folded contraction, or None when the basis does not have the second symmetry.

    On top of :func:`fold_operand` (n <-> N-n), a full-resolution basis (integer bins k = 0 .. N/2) satisfies
        cos(2 pi (N/2 - k) n / N) = (-1)^n cos(2 pi k n / N),   sin(2 pi (N/2 - k) n / N) = -(-1)^n sin(2 pi k n / N)
    for ANY window, so with the folded sums split by the parity of n,
        Ce[k] = sum_{n even} Bc[k][n] e[n]   Co[k] = sum_{n odd} Bc[k][n] e[n]
        Se[k] = sum_{n even} Bs[k][n] o[n]   So[k] = sum_{n odd} Bs[k][n] o[n]          (n = 1 .. N/2)
    bin k has re = Ce + Co, im = Se + So and bin N/2 - k has re = Ce - Co, im = -(Se - So): one radix-2 decimation
    step of the FFT.  The contraction then needs k = 1 .. N/4 only -- four chains of length N/4 over N/4 rows, HALF the
    multiply-adds of the once-folded form.  Bins 0 and N/2 are not produced (the caller checks that nothing reads
    them), and the n = 0 term must vanish (w0 == 0: every window that starts at zero, e.g. periodic Hann).

    Returns dict(basis_hi, basis_lo: float16 [4 * n_k, N/4] planes, chains Ce | Co | Se | So, row r <-> k = r + 1;
    columns of a chain ordered by increasing n of its parity; n_k = N/4; scale_inv).  The matching frame planes
    carry the even-n columns first, then the odd-n ones (``rvb_fold_split2_f16``).

    ``centre_doubled=True``: the operand of the contraction that folds in-kernel (``rvb_stft_mel_fused_pcm16``).  Its
    converter forms e[n] = p[n] + p[N-n] for EVERY column, so the centre sample n = N/2 (its own partner; last column
    of the even-n chain) arrives as 2 p[N/2]: that column of the cos rows carries half the weight (exact: a power of
    two).  The sin rows are zero there either way."""
    F, N = wcos.shape
    if N % 512 != 0 or N > 2048 or F != N // 2 + 1:
        return None
    half, quarter = N // 2, N // 4
    if quarter % FOLD2_TILE_K != 0:
        return None
    first = fold_operand(wcos, wsin, tol=tol, operand="tf32")
    if first is None or first["w0"] != 0.0:
        return None
    c64, s64 = wcos.astype(np.float64), wsin.astype(np.float64)
    bc = np.empty((F, half), np.float64)               # column c <-> n = c + 1 (as fold_operand, float64 mirror average)
    bs = np.zeros((F, half), np.float64)
    bc[:, :half - 1] = 0.5 * (c64[:, 1:half] + c64[:, N - 1:half:-1])
    bc[:, half - 1] = c64[:, half]
    bs[:, :half - 1] = 0.5 * (s64[:, 1:half] - s64[:, N - 1:half:-1])
    n = np.arange(1, half + 1)
    sgn = np.where(n % 2 == 0, 1.0, -1.0)              # (-1)^n
    k = np.arange(1, quarter + 1)
    scale = max(float(np.abs(c64).max()), 1e-30)
    if np.abs(bc[half - k] - sgn * bc[k]).max() > tol * scale or np.abs(bs[half - k] + sgn * bs[k]).max() > tol * scale:
        return None
    C = 0.5 * (bc[k] + sgn * bc[half - k])             # float64 average of the two mirror rows (<= 1 fp32 ulp apart)
    S = 0.5 * (bs[k] - sgn * bs[half - k])
    even, odd = np.flatnonzero(n % 2 == 0), np.flatnonzero(n % 2 == 1)
    mat = np.concatenate([C[:, even], C[:, odd], S[:, even], S[:, odd]], axis=0)     # [4 * n_k, N/4]
    if centre_doubled:
        assert n[even[-1]] == half
        mat[:quarter, quarter - 1] *= 0.5
    hi, lo, scale_inv = f16_split64(mat)
    return dict(basis_hi=hi, basis_lo=lo, n_k=quarter, scale_inv=scale_inv, even_cols=even, odd_cols=odd)


def mel_epilogue_table2(mel_basis, n_fft, chunk=EPILOGUE_CHUNK_BINS):
    """Tables for the Mel epilogue of the twice-folded contraction: float32 [2 * n_k, 4], n_k = n_fft / 4.

    Rows [0, n_k): the ASCENDING stream, row q <-> bin q + 1: (w0, w1, band0) as in :func:`mel_epilogue_table`.
    Rows [n_k, 2 n_k): the MIRRORED stream, row q <-> bin n_fft/2 - 1 - q, walked downwards in frequency.  In
    reversed band coordinates band' = n_mels - 1 - band the walk is again non-decreasing, so the same rotating
    accumulators serve: the row holds (w1, w0, band0' = n_mels - 2 - band0) -- it adds w1 P to band' band0' and w0 P
    to band0' + 1.  The mirrored row of bin n_fft/4 (q = n_k - 1) is zero: the ascending stream owns that bin.
    None when the bank cannot be represented: see :func:`mel_epilogue_table`; in addition bins 0 and n_fft/2 must
    carry no weight."""
    mb = np.asarray(mel_basis, dtype=np.float32)
    n_mels, F = mb.shape
    half, n_k = n_fft // 2, n_fft // 4
    if F != half + 1 or np.any(mb[:, 0] != 0) or np.any(mb[:, half] != 0):
        return None
    try:
        band0, w0, w1, k_begin, k_end = banded_filterbank(mb)
    except ValueError:
        return None
    b = band0.copy()
    b[:k_begin] = band0[k_begin]
    b[k_end:] = band0[k_end - 1]
    # bit-reproducibility: a band may receive at most two partial sums.  Streams are cut into chunks of `chunk` rows.
    row_of = np.empty(F, np.int64)
    row_of[1:n_k + 1] = np.arange(n_k)                                  # ascending rows
    row_of[n_k + 1:half] = n_k + (half - 1 - np.arange(n_k + 1, half))  # mirrored rows
    for m in range(n_mels):
        nz = np.flatnonzero(mb[m])
        if len(nz) and len(np.unique(row_of[nz] // chunk)) > 2:
            return None
    tab = np.zeros((2 * n_k, 4), np.float32)
    bins_up = np.arange(1, n_k + 1)
    tab[:n_k, 0], tab[:n_k, 1] = w0[bins_up], w1[bins_up]
    tab[:n_k, 2] = b[bins_up].astype(np.int32).view(np.float32)
    bins_dn = half - 1 - np.arange(n_k)
    tab[n_k:, 0], tab[n_k:, 1] = w1[bins_dn], w0[bins_dn]
    tab[2 * n_k - 1, :2] = 0.0                                          # bin n_fft/4 belongs to the ascending stream
    bd = (n_mels - 2 - b[bins_dn]).astype(np.int32)
    assert np.all(np.diff(b[bins_up]) >= 0) and np.all(np.diff(bd) >= 0)
    tab[n_k:, 2] = bd.view(np.float32)
    return tab
