/*
 * rvb.h -- C ABI of librvb.so, the sm_100a (B200) kernels behind reconvat_b200.
 *
 * The reference (KinWaiCheuk/ReconVAT) is 100 % Python and has no FFI of its
 * own: its hot path is a sequence of ATen calls.  Each entry point below
 * replaces the ATen call sequence cited next to it (paths relative to the
 * reference tree); reconvat_b200/_lib.py binds them with ctypes and
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to float32 unless stated otherwise;
 *     all memory is owned by the caller (PyTorch tensors); the library keeps
 *     no global state besides a cache of TMA descriptors keyed by
 *     (pointer, shape);
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work
 *     on it (no host synchronisation);
 *   - return value: 0 = enqueued; negative = error (RVB_ERR_*), message via
 *     rvb_last_error() (thread-local).  Nothing throws across the boundary;
 *   - device-side anomalies (NaN/Inf in r_adv) are reported through a
 *     caller-owned device int (`status_flag`), never by aborting the kernel.
 */
#ifndef RVB_H_
#define RVB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVB_ABI_VERSION 1

#define RVB_OK 0
#define RVB_ERR_ARG (-1)    /* bad argument (shape / alignment / unsupported combination) */
#define RVB_ERR_CUDA (-2)   /* a CUDA runtime / driver call failed */
#define RVB_ERR_LAUNCH (-3) /* kernel launch failed */

typedef void* rvb_stream_t;

int rvb_abi_version(void);
const char* rvb_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py: gpu_launches) */
int64_t rvb_launch_count(void);

/* ------------------------------------------------------------------ front-end */

/* pad modes of STFT.forward, model/Spectrogram.py:209-218 */
#define RVB_PAD_REFLECT 0  /* center=True, pad_mode='reflect'  (nn.ReflectionPad1d(n_fft//2)) */
#define RVB_PAD_CONSTANT 1 /* center=True, pad_mode='constant' (nn.ConstantPad1d(n_fft//2, 0)) */
#define RVB_PAD_NONE 2     /* center=False */

/*
 * K0  pad + hop-blocking + tf32 hi/lo split.
 * Replaces: the caller-side trim `audio[:, :-1]` (model/self_attention_VAT.py:1100,1112,1296 --
 * express it through n_samples / audio_ld) and `padding(x)` (model/Spectrogram.py:216-218).
 * Writes two planes [n_seg][rows_per_seg][hop]: sig_hi = tf32(p), sig_lo = tf32(p - sig_hi),
 * where p is the padded signal laid out in non-overlapping hop-sized rows, so that sample n of
 * frame t is plane[(t + n / hop) * hop + n % hop]; positions past the padded length are zero.
 */
int rvb_pad_split(const float* audio, int64_t audio_ld, int n_seg, int n_samples, int pad, int pad_mode,
                  float* sig_hi, float* sig_lo, int rows_per_seg, int hop, rvb_stream_t stream);

/* epilogues of the STFT contraction, model/Spectrogram.py:226-237 and :458 */
#define RVB_EPI_POWER 0     /* (sqrt(re^2+im^2))^2   out0[b][k][t]     (:227,231 then **2.0 at :458) */
#define RVB_EPI_MAGNITUDE 1 /* sqrt(re^2+im^2)       out0[b][k][t]     (:227,231) */
#define RVB_EPI_COMPLEX 2   /* (re, -im)             out0[b][k][t][2]  (:234) */
#define RVB_EPI_PHASE 3     /* atan2(-im + 0.0, re)  out0[b][k][t]     (:237) */
#define RVB_EPI_POWER_P 4   /* sqrt(re^2+im^2)^power out0[b][k][t]     (general `power`, :458) */
/* OR-ed onto a single-float epilogue: write out0[b][t][k] (time-major, what rvb_mel_project reads) */
#define RVB_EPI_TIME_MAJOR 0x10

/*
 * K1  STFT as a dense contraction on tcgen05 tensor cores, 3xTF32 (hi*hi + hi*lo + lo*hi, fp32
 * accumulators in TMEM), TMA-staged operands.
 * Replaces: `conv1d(x, wsin, stride)` + `conv1d(x, wcos, stride)` (model/Spectrogram.py:219-220),
 * the magnitude (:226-231) and `** self.power` (:458) or the Complex / Phase formats (:234,237).
 *   sig_hi/lo   planes from rvb_pad_split
 *   basis_hi/lo [n_basis_rows][n_fft] row-major, tf32 hi/lo split of the windowed Fourier basis,
 *               rows grouped in tiles of 256: rows [256j, 256j+128) = wcos bins [128j, 128j+128),
 *               rows [256j+128, 256j+256) = wsin of the same bins  (n_basis_rows % 256 == 0)
 *   out0        n_out_bins rows per segment; the kernel writes bins [0, min(n_out_bins, n_basis_rows/2))
 * Constraints: n_fft % 32 == 0, hop % 32 == 0, all pointers 128-byte aligned.
 */
int rvb_stft_gemm(const float* sig_hi, const float* sig_lo, int n_seg, int rows_per_seg, int hop, int n_frames,
                  const float* basis_hi, const float* basis_lo, int n_basis_rows, int n_fft, int epilogue,
                  float power, float* out0, int n_out_bins, rvb_stream_t stream);

/*
 * K0f  pad + frame + FOLD + tf32 hi/lo split, for windows that are symmetric about n_fft/2
 * (w[n] == w[n_fft-n]; true for every periodic scipy window at win_length == n_fft, checked on the host).
 * For a real frame p[0..N) the windowed Fourier basis satisfies wcos[k][N-n] = wcos[k][n] and
 * wsin[k][N-n] = -wsin[k][n], so
 *     re[k] = w[0] p[0] + sum_{n=1}^{N/2} wcos[k][n] e[n],   e[n] = p[n] + p[N-n]  (e[N/2] = p[N/2])
 *     im[k] =             sum_{n=1}^{N/2-1} wsin[k][n] o[n], o[n] = p[n] - p[N-n]
 * which halves the contraction length (model/Spectrogram.py:219-220 computes the same sums unfolded).
 * Writes the operand planes a_hi/a_lo as [2][n_seg*n_frames][n_fft/2]: plane 0 = e, plane 1 = o
 * (column c <-> n = c+1; o's last column is 0), each split as hi = tf32(v), lo = tf32(v - hi) with the
 * rounding error of the fp32 add recovered (TwoSum), and optionally p0[frame] = p[0] (needed when w[0] != 0).
 */
int rvb_fold_split(const float* audio, int64_t audio_ld, int n_seg, int n_samples, int pad, int pad_mode, int n_fft,
                   int hop, int n_frames, float* a_hi, float* a_lo, float* p0, rvb_stream_t stream);

/*
 * K1f  the folded STFT contraction on tcgen05 (two K = n_fft/2 chains per tile: e x cos -> re, o x sin -> im).
 *   a_hi/a_lo     planes from rvb_fold_split
 *   basis_hi/lo   [2][n_bins_pad][n_fft/2] row-major: plane 0 = folded cos rows, plane 1 = folded sin rows
 *                 (n_bins_pad % 128 == 0)
 *   p0, w0        optional rank-1 term re += w0 * p0[frame] (pass NULL / 0 when the window starts at 0)
 * Same epilogues and output indexing as rvb_stft_gemm.
 */
int rvb_stft_gemm_folded(const float* a_hi, const float* a_lo, int n_seg, int n_frames, int n_fft,
                         const float* basis_hi, const float* basis_lo, int n_bins_pad, const float* p0, float w0,
                         int epilogue, float power, float* out0, int n_out_bins, rvb_stream_t stream);

/*
 * K1fb  one frequency bin from the folded planes in plain fp32 FMA (Nyquist bin of the STFT module).
 * wc_row / ws_row: that bin's folded fp32 basis rows (n_fft/2 floats each).
 */
int rvb_stft_bin_folded(const float* a_hi, const float* a_lo, int n_seg, int n_frames, int n_fft,
                        const float* wc_row, const float* ws_row, const float* p0, float w0, int bin, int epilogue,
                        float power, float* out0, int n_out_bins, rvb_stream_t stream);

/*
 * K0h / K1h / K1hb  the fp16 flavour of the folded contraction (3xFP16: hi*hi + hi*lo + lo*hi with fp32
 * accumulation in TMEM).  fp16 has tf32's 11 significant bits, so the split carries the same 22 operand bits as
 * 3xTF32 while `tcgen05.mma.kind::f16` runs at twice the `kind::tf32` rate.  Its 5-bit exponent is handled by
 * power-of-two block scaling: rvb_fold_split_f16 scales each frame row so that its largest |e|,|o| lies in
 * [2^14, 2^15) and writes row_scale_inv[frame] = 2^-s; the host scales the basis the same way
 * (basis_scale_inv); the epilogue multiplies the accumulator by row_scale_inv * basis_scale_inv (exact).
 *   a_hi/a_lo    IEEE binary16 planes [2][n_seg*n_frames][n_fft/2] (plane 0 = e, plane 1 = o), 128-byte aligned
 *   basis_hi/lo  IEEE binary16 [2][n_bins_pad][n_fft/2]
 * Replaces the same reference lines as rvb_fold_split / rvb_stft_gemm_folded / rvb_stft_bin_folded
 * (model/Spectrogram.py:209-231, :458).  Constraint: n_fft % 128 == 0.
 * Accumulation order (rvb_stft_gemm_folded_f16, rvb_stft_mel_folded_f16): hi*lo + lo*hi over the WHOLE contraction
 * first, hi*hi on top -- tcgen05.mma truncates when it adds into the fp32 accumulator, and this order keeps two of
 * the three truncations per 16 terms away from the large partial sums (log-Mel <= 3.2e-5 of float64 on every stress
 * signal).  The results are bit-identical to the CPU model of the tensor core's accumulation
 * (tests/test_gpu_frontend.py) run over the same planes.
 */
int rvb_fold_split_f16(const float* audio, int64_t audio_ld, int n_seg, int n_samples, int pad, int pad_mode,
                       int n_fft, int hop, int n_frames, void* a_hi, void* a_lo, float* row_scale_inv, float* p0,
                       rvb_stream_t stream);
/* PCM16 input as the dataset stores it: sample = pcm * gain with gain = 1/32768 (model/dataset.py:62,
 * `audio.float().div_(32768.0)`; exact in fp32).  Halves the host->device bytes of the front-end. */
int rvb_fold_split_f16_pcm16(const int16_t* audio, int64_t audio_ld, float gain, int n_seg, int n_samples, int pad,
                             int pad_mode, int n_fft, int hop, int n_frames, void* a_hi, void* a_lo,
                             float* row_scale_inv, float* p0, rvb_stream_t stream);
int rvb_stft_gemm_folded_f16(const void* a_hi, const void* a_lo, const float* row_scale_inv, int n_seg, int n_frames,
                             int n_fft, const void* basis_hi, const void* basis_lo, float basis_scale_inv,
                             int n_bins_pad, const float* p0, float w0, int epilogue, float power, float* out0,
                             int n_out_bins, rvb_stream_t stream);
int rvb_stft_bin_folded_f16(const void* a_hi, const void* a_lo, const float* row_scale_inv, int n_seg, int n_frames,
                            int n_fft, const float* wc_row, const float* ws_row, const float* p0, float w0, int bin,
                            int epilogue, float power, float* out0, int n_out_bins, rvb_stream_t stream);

/*
 * K1b  one frequency bin in plain fp32 FMA (used for the Nyquist bin, which would otherwise cost a
 * whole extra 256-column tile).  Same epilogues / output indexing as rvb_stft_gemm; `wcos_row`,
 * `wsin_row` are the fp32 windowed basis rows of that bin (model/Spectrogram.py:162-164).
 */
int rvb_stft_bin(const float* sig_hi, const float* sig_lo, int n_seg, int rows_per_seg, int hop, int n_frames,
                 const float* wcos_row, const float* wsin_row, int n_fft, int bin, int epilogue, float power,
                 float* out0, int n_out_bins, rvb_stream_t stream);

/*
 * K1m  the folded fp16 contraction with the Mel projection fused into its epilogue: the power (or magnitude /
 * power_p) spectrum never leaves the SM.  Replaces model/Spectrogram.py:219-231, :458 and `torch.matmul(self.mel_basis,
 * spec)` (:460) for filterbanks in which every bin feeds at most two adjacent bands (all triangular banks).
 *   spectrum   RVB_EPI_POWER | RVB_EPI_MAGNITUDE | RVB_EPI_POWER_P
 *   mel_tab    HOST pointer, [n_bins_pad][4] float: (w0, w1, band0 as int32 bits, 0): bin k adds w0 P[k] to band
 *              band0[k] and w1 P[k] to band band0[k]+1; band0 non-decreasing in k; n_bins_pad <= 1024.  The table is
 *              read during the call and travels as a 16 KB kernel parameter (constant bank, no shared-memory traffic)
 *   mel_out    [n_seg][n_mels][n_frames] (what MelSpectrogram.forward returns); zeroed by the call, then accumulated
 *              with RED.ADD -- at most two partial sums per element when no band straddles more than two 128-bin tiles
 * K2m  rvb_logmel_minmax: per-segment (min, max) keys of log(mel + log_offset) (model/self_attention_VAT.py:1102,
 * utils.py:96-97);  rvb_logmel_transpose: out[b][t][m] = (log(mel[b][m][t] + log_offset) - min) / (max - min)
 * (utils.py:100 and the .transpose(-1,-2) of self_attention_VAT.py:1104; minmax NULL: no normalisation;
 * log_offset < 0: no log).
 */
int rvb_stft_mel_folded_f16(const void* a_hi, const void* a_lo, const float* row_scale_inv, int n_seg, int n_frames,
                            int n_fft, const void* basis_hi, const void* basis_lo, float basis_scale_inv, int n_bins_pad,
                            const float* p0, float w0, int spectrum, float power, const float* mel_tab, int n_mels,
                            float* mel_out, rvb_stream_t stream);
/*
 * K0q / K1q  the TWICE-folded contraction (same references as K0h / K1m).  For integer bins the basis has a second
 * symmetry, cos(2 pi (N/2-k) n/N) = (-1)^n cos(2 pi k n/N), sin likewise with a minus sign: splitting the folded sums by
 * the parity of n yields bin k and bin N/2-k from the same four partial sums (one radix-2 decimation step), so the
 * contraction runs over k = 1 .. N/4 only -- half the multiply-adds of K1m.  Bins 0 and N/2 are not produced.
 * n_fft: a multiple of 512, at most 2048.
 *   rvb_fold_split2_f16[_pcm16]: as rvb_fold_split_f16[_pcm16], but a row's columns are ordered even n first
 *     (n = 2, 4, .., N/2), then odd n (n = 1, 3, .., N/2-1); no p0 output (the n = 0 term must vanish: w0 == 0).
 *   rvb_stft_mel_folded2_f16: power spectrum + Mel projection.
 *     basis_hi/lo  [4 * N/4][N/4] fp16: chains Ce | Co | Se | So, row r <-> k = r + 1, columns in increasing n of the
 *                  chain's parity (basis.fold2_operand), block-scaled by 2^s (basis_scale_inv = 2^-s)
 *     mel_tab      HOST pointer, [2 * N/4][4] float (basis.mel_epilogue_table2): rows [0, N/4) the ascending stream
 *                  (row q <-> bin q + 1), rows [N/4, N/2) the mirrored stream (row q <-> bin N/2 - 1 - q) in
 *                  reversed band coordinates; n_fft <= 2048
 *     mel_out      [2][n_seg][n_mels][n_frames]: the cos^2 and the sin^2 part of the Mel power spectrogram (the
 *                  projection is linear in P = re^2 + im^2, so a unit of the kernel handles ONE component: half the
 *                  L2 -> SM operand traffic per multiply-add); the consumer adds the planes (rvb_logmel_normalise
 *                  does).  Zeroed by the call, accumulated with RED.ADD, at most two partial sums per element
 */
int rvb_fold_split2_f16(const float* audio, int64_t audio_ld, int n_seg, int n_samples, int pad, int pad_mode,
                        int n_fft, int hop, int n_frames, void* a_hi, void* a_lo, float* row_scale_inv,
                        rvb_stream_t stream);
int rvb_fold_split2_f16_pcm16(const int16_t* audio, int64_t audio_ld, float gain, int n_seg, int n_samples, int pad,
                              int pad_mode, int n_fft, int hop, int n_frames, void* a_hi, void* a_lo,
                              float* row_scale_inv, rvb_stream_t stream);
int rvb_stft_mel_folded2_f16(const void* a_hi, const void* a_lo, const float* row_scale_inv, int n_seg, int n_frames,
                             int n_fft, const void* basis_hi, const void* basis_lo, float basis_scale_inv,
                             const float* mel_tab, int n_mels, float* mel_out, rvb_stream_t stream);
/*
 * K0x / K1x  the twice-folded contraction WITHOUT materialised frame planes (PCM16 input, the dataset's storage format,
 * model/dataset.py:19-62; same references as K0q / K1q: model/Spectrogram.py:209-231, :458, :460).  K0q writes every
 * sample four times as fp16 hi/lo of e and o (16 bytes per sample of the hop) and K1q reads that back; here
 *   rvb_pad_parity_pcm16: reflect-pads the signal (or constant / none, as rvb_fold_split*) and stores it ONCE, split by
 *     sample parity, in offset binary: planes[q][b][i] = padded sample 2 i + q of segment b, + 32768 (uint16; 32768
 *     past the end).  plane_len: rvb_parity_plane_len(...) elements (a multiple of 8), planes 128-byte aligned.
 *   rvb_stft_mel_fused_pcm16: converter warps inside the tcgen05 kernel read those planes, form e = p[n] + p[N-n] /
 *     o = p[n] - p[N-n] exactly, split them into fp16 hi + lo and write the swizzled operand tiles in shared memory;
 *     contraction, epilogue, basis layout, Mel table and mel_out exactly as rvb_stft_mel_folded2_f16, except that the
 *     basis carries HALF the weight in the centre column of the even-n cos chain (the centre sample is its own
 *     partner: e = 2 p[N/2]; basis.fold2_operand(centre_doubled=True)).  gain: the PCM scale (1/32768), a power of two.
 *     hop: a multiple of 16.
 */
int64_t rvb_parity_plane_len(int n_samples, int pad, int pad_mode, int n_fft, int hop, int n_frames);
int rvb_pad_parity_pcm16(const int16_t* audio, int64_t audio_ld, int n_seg, int n_samples, int pad, int pad_mode,
                         uint16_t* planes, int64_t plane_len, rvb_stream_t stream);
int rvb_stft_mel_fused_pcm16(const uint16_t* planes, int64_t plane_len, int n_seg, int n_frames, int n_fft, int hop,
                             float gain, const void* basis_hi, const void* basis_lo, float basis_scale_inv,
                             const float* mel_tab, int n_mels, float* mel_out, rvb_stream_t stream);
int rvb_logmel_minmax(const float* mel, int n_seg, int64_t n_per_seg, float log_offset, uint32_t* minmax,
                      rvb_stream_t stream);
int rvb_logmel_transpose(const float* mel, int n_seg, int n_mels, int n_frames, float log_offset,
                         const uint32_t* minmax, float* out, rvb_stream_t stream);
/*
 * K2m + K3m in one pass: out[b][t][m] = (log(mel[b][m][t] + log_offset) - min_b) / (max_b - min_b).  A thread-block
 * cluster of 8 CTAs owns one segment: its log-Mel values stay in shared memory between the min/max reduction
 * (exchanged through distributed shared memory) and the normalised, transposed write -- mel is read once.
 * Replaces `torch.log(spec + 1e-5)`, Normalization('imagewise').transform and `.transpose(-1,-2)`
 * (model/self_attention_VAT.py:1102-1104, model/utils.py:93-100).  Bit-identical to rvb_logmel_minmax +
 * rvb_logmel_transpose, which it falls back to when a segment does not fit the cluster's shared memory
 * (or RVB_NO_NORM_FUSION=1).  minmax: uint32 [n_seg][2], receives the keys (required).
 * mel_b: NULL, or the second plane of rvb_stft_mel_folded2_f16 -- the kernel then reads mel + mel_b (fused path only).
 */
int rvb_logmel_normalise(const float* mel, const float* mel_b, int n_seg, int n_mels, int n_frames, float log_offset,
                         uint32_t* minmax, float* out, rvb_stream_t stream);

#define RVB_LAYOUT_BINS_MAJOR 0 /* out[b][m][t]  -- what MelSpectrogram.forward returns (:460) */
#define RVB_LAYOUT_TIME_MAJOR 1 /* out[b][t][m]  -- after `.transpose(-1,-2)` (self_attention_VAT.py:1104) */

/*
 * K2  banded Mel projection (+ optional log compression and per-segment min/max).
 * Replaces: `torch.matmul(self.mel_basis, spec)` (model/Spectrogram.py:460),
 * `torch.log(spec + 1e-5)` (model/self_attention_VAT.py:1102), the two reductions of
 * Normalization('imagewise') (model/utils.py:96-97) and, with RVB_LAYOUT_TIME_MAJOR, the
 * `.transpose(-1,-2)` (model/self_attention_VAT.py:1104).
 * The filterbank is passed by rows: band m reads bins [band_lo[m], band_lo[m] + band_len[m]) with weights
 * band_w[j][m], j < band_len[m] (band_w is [max_len][n_mels], band fastest): any filterbank whose rows have
 * contiguous support.
 *   power     [n_seg * n_frames][n_bins], time-major (RVB_EPI_TIME_MAJOR epilogue of the contraction)
 *   log_offset < 0: no log;  >= 0: out = logf(mel + log_offset)
 *   minmax    NULL, or uint32 [n_seg][2] receiving order-preserving keys of (min, max) of `out`
 *             per segment (decoded by rvb_normalise); the call zeroes it first.
 */
int rvb_mel_project(const float* power, int n_seg, int n_frames, int n_bins, const int32_t* band_lo,
                    const int32_t* band_len, const float* band_w, int max_len, int n_mels, float log_offset,
                    int layout, float* out, uint32_t* minmax, rvb_stream_t stream);

/*
 * Per-segment min/max keys of an arbitrary [n_seg][n_per_seg] tensor
 * (model/utils.py:96-97 when Normalization is used on its own).
 */
int rvb_minmax(const float* x, int n_seg, int64_t n_per_seg, uint32_t* minmax, rvb_stream_t stream);

/*
 * K3  y = (x - min) / (max - min) per segment, in place when y == x
 * (model/utils.py:100; no epsilon: a constant image gives NaN exactly as the reference).
 */
int rvb_normalise(const float* x, float* y, int n_seg, int64_t n_per_seg, const uint32_t* minmax,
                  rvb_stream_t stream);

/*
 * K3f  Normalization('framewise') (model/utils.py:85-92): x, y are [n_seg][n_bins][n_frames]; per (segment, frame)
 * min / max over the bins, y = (x - min) / (max - min), NaN -> 0.  In place when y == x.
 */
int rvb_normalise_framewise(const float* x, float* y, int n_seg, int n_bins, int n_frames, rvb_stream_t stream);

/* ------------------------------------------------------------------ VAT loop */

/*
 * V1  x_adv = clamp(x + xi * d / ||d||_row, 0, 1)      rows of `row_len` (229) floats.
 * Replaces: `_l2_normalize(d)` + `XI * .` + `(x + r).clamp(0,1)`
 * (model/self_attention_VAT.py:176-177, :240-246; model/UNet_onset.py:130-131;
 *  model/onset_frame_VAT.py:182-183; model/VAT.py:27-28 with do_clamp = 0).
 */
int rvb_vat_perturb(const float* x, const float* d, float* x_adv, int64_t n_rows, int row_len, float xi,
                    int do_clamp, rvb_stream_t stream);
/*
 * V0 + V1  rvb_vat_perturb with the direction DRAWN IN THE KERNEL (model/self_attention_VAT.py:172, :176-177):
 * d = torch.randn_like(x) bit for bit -- Philox4x32-10 keyed by the generator's seed, Box-Muller, ATen's element <->
 * (thread, call, component) mapping for a contiguous tensor -- then x_adv as rvb_vat_perturb.  No ATen kernel, no
 * round trip of d through HBM for this step; d_out (nullable) receives d for rvb_vat_finalize.
 *   seed, offset    the CUDA generator's seed and Philox offset BEFORE the draw (offset % 4 == 0)
 *   aten_threads   256 * min(#SM * (maxThreadsPerSM / 256), ceil(numel / 256)): ATen's launch geometry
 *   increment       ((numel - 1) / (4 * aten_threads) + 1) * 4: what the caller adds to the generator's offset
 *   dev_state       NULL, or device uint64[3] = {seed, offset, 0}: read INSTEAD of (seed, offset), and the offset is
 *                   advanced by `increment` when the kernel ends -- a captured CUDA graph replays a continuing stream
 */
int rvb_vat_perturb_draw(const float* x, float* d_out, float* x_adv, int64_t n_rows, int row_len, float xi, int do_clamp,
                         uint64_t seed, uint64_t offset, uint32_t aten_threads, uint64_t increment, uint64_t* dev_state,
                         rvb_stream_t stream);
/*
 * V0  d = torch.randn_like(x) (model/self_attention_VAT.py:172) as a kernel of this library: the same bits as ATen's
 * normal_ kernel for a contiguous float tensor of n elements (arguments as rvb_vat_perturb_draw), one Philox call per
 * four elements like ATen, all trips of a thread issued at once.  For steps that run beside a tensor-bound kernel of
 * another stream, where the 4x Philox work of the fused draw costs more than a round trip of d through HBM.
 */
int rvb_randn_like(float* out, int64_t n, uint64_t seed, uint64_t offset, uint32_t aten_threads, uint64_t increment,
                   uint64_t* dev_state, rvb_stream_t stream);

/*
 * V2  grad = gscale * (p - y) / max((1 - p) * p, 1e-12) / n   (d mean-BCE / d p, ATen's formula).
 * Replaces: the backward of `F.binary_cross_entropy(y_pred, y_ref)`
 * (model/self_attention_VAT.py:182-183).  `gscale_dev` is an optional device scalar (upstream grad).
 */
int rvb_bce_grad(const float* p, const float* y, float* grad, int64_t n, const float* gscale_dev, float gscale,
                 rvb_stream_t stream);

/*
 * V3  power-iteration backward + finalisation, one pass over (g, d, x):
 *   n = ||d||, dhat = d/n, m = [0 <= x + xi*dhat <= 1], gd = xi * g * m
 *   d' = scale * (gd / n - d * sum(gd * d) / n^3)                 (== d.grad * scale)
 *   dhat' = d'/||d'||, r_adv = eps * dhat', x_adv = clamp(x + r_adv, 0, 1)
 * Replaces: autograd through clamp/add/mul/div/norm, `d.grad.detach()*1e10`,
 * `eps*_l2_normalize(d)`, the NaN asserts, `(x + r_adv).clamp(0,1)` and the recomputed
 * `_l2_normalize(d)` (model/self_attention_VAT.py:183-202 and siblings).
 * status_flag (device int, caller zeroes): bit0 = NaN in r_adv, bit1 = Inf in r_adv.
 */
int rvb_vat_finalize(const float* g, const float* d, const float* x, float* r_adv, float* x_adv, float* d_hat,
                     int64_t n_rows, int row_len, float xi, float eps, float scale, int do_clamp,
                     int32_t* status_flag, rvb_stream_t stream);

/*
 * V3b finalisation only (n_power == 0: the random direction is used as is):
 *   dhat = d/||d||, r_adv = eps*dhat, x_adv = clamp(x + r_adv)     (:188-194 with d = randn)
 */
int rvb_vat_direct(const float* d, const float* x, float* r_adv, float* x_adv, float* d_hat, int64_t n_rows,
                   int row_len, float eps, int do_clamp, int32_t* status_flag, rvb_stream_t stream);

/*
 * V3 / V3b with the step's by-products folded in: the same kernels, but the per-block NaN / Inf bits and sum |dhat|
 * go to a workspace and the last block to finish WRITES status_flag (no zeroing by the caller) and
 *   dhat_abs_mean = mean |dhat|     -- the `r_norm.abs().mean()` that run_on_batch logs after every VAT call
 *                                      (model/self_attention_VAT.py:1149-1150), summed in a fixed order in double.
 * g == NULL selects V3b (n_power == 0).  d_hat == NULL: the normalised direction itself is not stored (a caller that
 * only logs its mean saves a sixth of the kernel's traffic).  workspace: rvb_vat_stats_workspace_bytes(n_rows) bytes, zeroed once by the
 * caller (self-cleaning); one workspace per stream that may run the kernel concurrently.
 */
int64_t rvb_vat_stats_workspace_bytes(int64_t n_rows);
int rvb_vat_finalize_stats(const float* g, const float* d, const float* x, float* r_adv, float* x_adv, float* d_hat,
                           int64_t n_rows, int row_len, float xi, float eps, float scale, int do_clamp,
                           int32_t* status_flag, float* dhat_abs_mean, void* workspace, int64_t workspace_bytes,
                           rvb_stream_t stream);

/*
 * V4  loss = mean( -(y*max(log p,-100) + (1-y)*max(log1p(-p),-100)) ), deterministic two-level sum.
 * Replaces: `F.binary_cross_entropy(y_pred, y_ref)` forward (model/self_attention_VAT.py:200).
 * workspace: >= RVB_BCE_WORKSPACE_FLOATS floats, zero-initialised once by the caller (self-cleaning).
 */
#define RVB_BCE_WORKSPACE_FLOATS 1032
int rvb_bce_mean(const float* p, const float* y, int64_t n, float* loss, float* workspace, rvb_stream_t stream);

/*
 * V2' / V4'  the other divergences of the reference's VAT flavours, same kernels as V2 / V4 (workspace as V4):
 *   RVB_DIV_BCE  F.binary_cross_entropy(p, y)                        denom = n
 *   RVB_DIV_BKL  binary_kl_div(p, y) (model/self_attention_VAT.py:248-255, KL_Div=True): both clamped to
 *                [1e-4, 0.9999], F.kl_div(log [y, 1-y], [p, 1-p], reduction='batchmean')   denom = p.shape[0]
 *   RVB_DIV_MSE  F.mse_loss(p, y) (model/onset_frame_VAT.py:232, stepwise_VAT_frame_stack)  denom = n
 * loss = sum / denom;  grad = gscale * gscale_dev[0] * d(sum)/dp / denom.
 */
#define RVB_DIV_BCE 0
#define RVB_DIV_BKL 1
#define RVB_DIV_MSE 2
int rvb_div_grad(int kind, const float* p, const float* y, float* grad, int64_t n, double denom,
                 const float* gscale_dev, float gscale, rvb_stream_t stream);
int rvb_div_mean(int kind, const float* p, const float* y, int64_t n, double denom, float* loss, float* workspace,
                 rvb_stream_t stream);

/*
 * V1b / V3b'  binwise=True flavour of _l2_normalize, d / (|d| + 1e-8) (model/self_attention_VAT.py:242-243;
 * stepwise_VAT(..., binwise=True)): elementwise, n = number of elements.  g == NULL: n_power == 0 (d is used as is).
 */
int rvb_vat_perturb_binwise(const float* x, const float* d, float* x_adv, int64_t n, float xi, int do_clamp,
                            rvb_stream_t stream);
int rvb_vat_finalize_binwise(const float* g, const float* d, const float* x, float* r_adv, float* x_adv, float* d_hat,
                             int64_t n, float xi, float eps, float scale, int do_clamp, int32_t* status_flag,
                             rvb_stream_t stream);

/* ------------------------------------------------------------------ caller-side: local-window attention */

/*
 * f2  the dense projections of the caller's sequence model on tensor cores: W_q / W_k / W_v of
 * MutliHeadAttention1D (nn.Linear, model/self_attention_VAT.py:54-56, applied :70-71) and their autograd backward.
 * PyTorch runs them as SIMT SGEMMs (TF32 is off by default for matmul); here they are 3xTF32 tcgen05 contractions
 * (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM: ~2^-21 relative, inside fp32 SGEMM's own rounding).
 *   rvb_split_tf32: x [rows][cols] (row stride ld) -> tf32 hi / lo planes.  transpose == 0: planes[r][offset + c];
 *     transpose == 1: planes[c][offset + r] (for contractions over x's rows: dW = dY^T X).  Plane row stride out_ld; the
 *     columns [offset + extent, zero_to) are zero-filled (pad the contraction length to a multiple of 32).
 *   rvb_gemm_nt_tf32x3: C[m][n] = A[m][k] . B[n][k]^T from such planes ([rows][k_pad], 128-byte aligned).  k_split > 1
 *     cuts the contraction into k_split slices of whole 32-element blocks (none empty); slice s writes its partial
 *     product to c + s * split_stride and the caller adds the partials (few output tiles, long contraction: dW).
 */
int rvb_split_tf32(const float* x, int64_t rows, int cols, int64_t ld, int transpose, float* hi, float* lo, int64_t out_ld,
                   int64_t offset, int64_t zero_to, rvb_stream_t stream);
int rvb_gemm_nt_tf32x3(const float* a_hi, const float* a_lo, int64_t m, const float* b_hi, const float* b_lo, int n,
                       int k_pad, float* c, int64_t ldc, int k_split, int64_t split_stride, rvb_stream_t stream);
/*
 * A1-A3  MutliHeadAttention1D.forward (model/self_attention_VAT.py:22-91; same class in model/UNet_onset.py:22,
 * model/self_attention.py:6) after the three nn.Linear projections, and its backward:
 *   energy[b,l,h,w] = sum_c q[b,l,h,c] k[b,l+w-P,h,c] + bias[b,l,h,w],  P = (W-1)/2, rows outside [0,L) are zero
 *   att = softmax_w(energy);  out[b,l,h,c] = sum_w att[b,l,h,w] v[b,l+w-P,h,c]
 * Replaces F.pad / unfold / (q*k).sum / softmax / (att*v).sum (:64-88) without materialising the (B, L, C, W) unfolded
 * k and v.  The relative-position term of :76 does not depend on k: bias = q . rel (and dq += dE . rel^T,
 * d rel = q^T dE in the backward) are plain batched GEMMs left to the caller; bias may be NULL (position=False).
 * q, k, v, out, dq, dk, dv, dout: [B][L][G*D];  att, dE, bias: [B][L][G][W].  W odd, <= 32;  D <= 508.
 *   rvb_local_attn_bwd_q   dE = softmax backward of (dout . v), dq = dE . k
 *   rvb_local_attn_bwd_kv  dk[m] = sum_w dE[m-w+P][w] q[m-w+P],  dv[m] = sum_w att[m-w+P][w] dout[m-w+P]
 */
int rvb_local_attn_fwd(const float* q, const float* k, const float* v, const float* bias, int B, int L, int G, int D,
                       int W, float* out, float* att, rvb_stream_t stream);
int rvb_local_attn_bwd_q(const float* dout, const float* att, const float* k, const float* v, int B, int L, int G,
                         int D, int W, float* dE, float* dq, rvb_stream_t stream);
int rvb_local_attn_bwd_kv(const float* q, const float* dout, const float* att, const float* dE, int B, int L, int G,
                          int D, int W, float* dk, float* dv, rvb_stream_t stream);

/*
 * D1  note decoding (model/decoding.py:4-55, extract_notes_wo_velocity; driven by transcribe_files.py:12-40):
 * onsets, frames: [n_frames][n_pitches] posteriors.  start[t][p] = 1 where a note begins (thresholded onset roll rises;
 * rule1: and the frame roll is on), offset[t][p] = first frame u >= t where neither thresholded roll is on (n_frames
 * if none).  The note list is then nonzero(start) with intervals (t, offset[t][p]) -- integer work, bit-exact.
 */
int rvb_note_offsets(const float* onsets, const float* frames, int n_frames, int n_pitches, float onset_threshold,
                     float frame_threshold, int rule1, uint8_t* start, int32_t* offset, rvb_stream_t stream);

/*
 * B1-B4  nn.BatchNorm2d of the caller's U-Net (model/self_attention_VAT.py:848-850, :865-869; the same blocks in
 * model/UNet_onset.py:190-211), fp32 NCHW, forward and backward (consumer side of the hot path).  Replaces
 * aten::cudnn_batch_norm / aten::cudnn_batch_norm_backward (train) and the eval-mode native_batch_norm.
 *   x, y, dy, dx   [n][c][hw] contiguous
 *   splits         slices per channel, from rvb_bn_splits(n, c, hw) (the same value in the calls of one pass)
 *   partials       float64 [c][splits][2] workspace
 *   rvb_bn_reduce    dy == NULL: partial sums of (x - K), (x - K)^2 with K = x[0][ch][0] (shifted: no cancellation);
 *                    else partial sums of dy, dy (x - mean[ch])
 *   rvb_bn_forward   mean, biased variance from the partials; y = (x - mean) / sqrt(var + eps) * gamma + beta;
 *                    save_mean / save_invstd for the backward; running_mean / running_var (NULL: not tracked) updated
 *                    with `momentum` and the UNBIASED variance, as nn.BatchNorm2d does.  gamma / beta NULL: 1 / 0
 *   rvb_bn_apply     eval mode: the same formula with the given mean / invstd
 *   rvb_bn_backward  training: dx = (dy - mean(dy) - xhat mean(dy xhat)) invstd gamma; eval (training = 0):
 *                    dx = dy invstd gamma; dgamma = sum(dy xhat), dbeta = sum(dy); any of dx / dgamma / dbeta may be NULL
 */
int rvb_bn_splits(int n, int c, int64_t hw);
/* rvb_bn_train_forward = rvb_bn_reduce + rvb_bn_forward, rvb_bn_train_backward = rvb_bn_reduce + rvb_bn_backward with
 * splits = rvb_bn_splits(n, c, hw), in ONE host call; partials must hold c * 64 * 2 doubles. */
int rvb_bn_train_forward(const float* x, int n, int c, int64_t hw, const float* gamma, const float* beta, float eps,
                         float momentum, float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                         float* y, double* partials, rvb_stream_t stream);
int rvb_bn_train_backward(const float* x, const float* dy, int n, int c, int64_t hw, const float* gamma, const float* mean,
                          const float* invstd, int training, float* dx, float* dgamma, float* dbeta, double* partials,
                          rvb_stream_t stream);
/* The same three operations for torch.channels_last tensors: x, y, dy, dx in [n][h][w][c] physical order, pixels = n h w,
 * c % 4 == 0, c <= 1024, all pointers 16-byte aligned.  A block owns a range of pixels and all channels, a thread one
 * group of four channels; the last block of the reduction (ticket) turns the per-block partial sums into the channel
 * statistics.  workspace: rvb_bn_nhwc_workspace_bytes(c) bytes whose LAST 16 bytes (the ticket) are zero before the
 * first use; the kernels leave them zero. */
int64_t rvb_bn_nhwc_workspace_bytes(int c);
int rvb_bn_train_forward_nhwc(const float* x, int64_t pixels, int c, const float* gamma, const float* beta, float eps,
                              float momentum, float* running_mean, float* running_var, float* save_mean,
                              float* save_invstd, float* y, void* workspace, rvb_stream_t stream);
int rvb_bn_apply_nhwc(const float* x, int64_t pixels, int c, const float* mean, const float* invstd, const float* gamma,
                      const float* beta, float* y, rvb_stream_t stream);
int rvb_bn_train_backward_nhwc(const float* x, const float* dy, int64_t pixels, int c, const float* gamma,
                               const float* mean, const float* invstd, int training, float* dx, float* dgamma,
                               float* dbeta, void* workspace, rvb_stream_t stream);
int rvb_bn_reduce(const float* x, const float* dy, const float* mean, int n, int c, int64_t hw, int splits,
                  double* partials, rvb_stream_t stream);
int rvb_bn_forward(const float* x, int n, int c, int64_t hw, int splits, const double* partials, const float* gamma,
                   const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                   float* save_mean, float* save_invstd, float* y, rvb_stream_t stream);
int rvb_bn_apply(const float* x, int n, int c, int64_t hw, const float* mean, const float* invstd, const float* gamma,
                 const float* beta, float* y, rvb_stream_t stream);
int rvb_bn_backward(const float* x, const float* dy, int n, int c, int64_t hw, int splits, const double* partials,
                    const float* gamma, const float* mean, const float* invstd, int training, float* dx, float* dgamma,
                    float* dbeta, rvb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RVB_H_ */
