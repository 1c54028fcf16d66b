// HBM-bound front-end kernels around the STFT contraction:
//   K0 pad + hop-blocking + tf32 split     (model/Spectrogram.py:209-218)
//   K1b single frequency bin in fp32       (Nyquist bin of model/Spectrogram.py:219-220)
//   K2 banded Mel + log + min/max          (model/Spectrogram.py:460, self_attention_VAT.py:1102, utils.py:96-97)
//   K3 imagewise normalise                 (model/utils.py:100)
#include <cooperative_groups.h>

#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "rvb_common.cuh"

namespace rvb {

extern void count_launch();

// ------------------------------------------------------------------ K0
// One thread produces 4 consecutive plane samples (float4 stores; float4 loads in the interior).
template <typename TIn>
__device__ __forceinline__ float padded_sample(const TIn* __restrict__ a, int64_t i, int n, int pad, int mode) {
  // i: index into the padded signal of length n + 2*pad (or n when mode == NONE)
  int64_t j = i - pad;
  if (mode == RVB_PAD_NONE) j = i;
  if (j < 0) {
    if (mode != RVB_PAD_REFLECT) return 0.f;
    j = -j;                                   // ReflectionPad1d: edge sample not repeated
  } else if (j >= n) {
    if (mode != RVB_PAD_REFLECT) return 0.f;
    j = 2 * (int64_t)(n - 1) - j;
  }
  return (float)__ldg(a + j);
}

__global__ void __launch_bounds__(256)
pad_split_kernel(const float* __restrict__ audio, int64_t audio_ld, int n_samples, int pad, int mode,
                 float* __restrict__ sig_hi, float* __restrict__ sig_lo, int64_t plane_per_seg) {
  const int b = blockIdx.y;
  const float* a = audio + (int64_t)b * audio_ld;
  const int64_t padded = (mode == RVB_PAD_NONE) ? n_samples : (int64_t)n_samples + 2 * pad;
  const int64_t off = (mode == RVB_PAD_NONE) ? 0 : pad;
  const bool src_vec = ((reinterpret_cast<uintptr_t>(a) & 15u) == 0) && ((off & 3) == 0);
  for (int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i4 < plane_per_seg;
       i4 += (int64_t)gridDim.x * blockDim.x * 4) {
    float v[4];
    const int64_t j0 = i4 - off;
    if (src_vec && j0 >= 0 && j0 + 3 < n_samples) {
      float4 t = __ldg(reinterpret_cast<const float4*>(a + j0));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int64_t i = i4 + e;
        v[e] = (i < padded) ? padded_sample(a, i, n_samples, pad, mode) : 0.f;
      }
    }
    float4 hi, lo;
    hi.x = to_tf32(v[0]); lo.x = to_tf32(v[0] - hi.x);
    hi.y = to_tf32(v[1]); lo.y = to_tf32(v[1] - hi.y);
    hi.z = to_tf32(v[2]); lo.z = to_tf32(v[2] - hi.z);
    hi.w = to_tf32(v[3]); lo.w = to_tf32(v[3] - hi.w);
    const int64_t o = (int64_t)b * plane_per_seg + i4;
    *reinterpret_cast<float4*>(sig_hi + o) = hi;
    *reinterpret_cast<float4*>(sig_lo + o) = lo;
  }
}

// ------------------------------------------------------------------ K0f
// One block per frame (grid-stride).  The frame's n_fft samples are staged in smem (float4 loads for
// interior frames, reflect-aware scalar loads at the segment edges), shifted by 3 floats so that the forward
// run p[c+1 .. c+4] of a 4-column group is one aligned LDS.128.  Each thread then emits 4 consecutive columns
// of the e and o planes as 16-byte stores.  Each audio sample is read by the 4 frames that overlap it: L2 hits.
// The fp32 add/sub rounds at 2^-24, below the 2^-22 of the tf32 split (and exact for int16-derived audio).
__device__ __forceinline__ void split2(float v, float& hi, float& lo) {
  hi = to_tf32(v);
  lo = to_tf32(v - hi);
}

__global__ void __launch_bounds__(256)
fold_split_kernel(const float* __restrict__ audio, int64_t audio_ld, int n_seg, int n_samples, int pad, int mode,
                  int n_fft, int hop, int n_frames, float* __restrict__ a_hi, float* __restrict__ a_lo,
                  float* __restrict__ p0) {
  extern __shared__ __align__(16) float frame_s[];         // frame_s[i + 3] = p[i], i in [0, n_fft)
  float* frame = frame_s + 3;
  const int half = n_fft >> 1;
  const int64_t n_rows = (int64_t)n_seg * n_frames;
  const int64_t padded = (mode == RVB_PAD_NONE) ? n_samples : (int64_t)n_samples + 2 * pad;
  const int off = (mode == RVB_PAD_NONE) ? 0 : pad;
  for (int64_t f = blockIdx.x; f < n_rows; f += gridDim.x) {
    const int b = (int)(f / n_frames), t = (int)(f - (int64_t)b * n_frames);
    const float* a = audio + (int64_t)b * audio_ld;
    const int64_t start = (int64_t)t * hop;                // index of p[0] in the padded signal
    const int64_t j0 = start - off;                        // ... and in the audio row
    __syncthreads();                                       // previous iteration's readers are done
    const bool interior = j0 >= 0 && j0 + n_fft <= n_samples && ((j0 & 3) == 0) &&
                          ((reinterpret_cast<uintptr_t>(a) & 15u) == 0);
    if (interior) {
      const float4* src = reinterpret_cast<const float4*>(a + j0);
      for (int i = threadIdx.x; i < (n_fft >> 2); i += blockDim.x) {
        const float4 v = __ldg(src + i);
        frame[4 * i + 0] = v.x; frame[4 * i + 1] = v.y; frame[4 * i + 2] = v.z; frame[4 * i + 3] = v.w;
      }
    } else {
      for (int i = threadIdx.x; i < n_fft; i += blockDim.x) {
        const int64_t pi = start + i;
        frame[i] = (pi < padded) ? padded_sample(a, pi, n_samples, pad, mode) : 0.f;
      }
    }
    __syncthreads();
    float4* e_hi = reinterpret_cast<float4*>(a_hi + f * half);
    float4* e_lo = reinterpret_cast<float4*>(a_lo + f * half);
    float4* o_hi = reinterpret_cast<float4*>(a_hi + (n_rows + f) * half);
    float4* o_lo = reinterpret_cast<float4*>(a_lo + (n_rows + f) * half);
    for (int q = threadIdx.x; q < (half >> 2); q += blockDim.x) {
      const int c = q << 2;                                // columns c .. c+3  <->  n = c+1 .. c+4
      const float4 x = *reinterpret_cast<const float4*>(frame_s + c + 4);        // p[c+1 .. c+4]
      float y[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) y[i] = frame[n_fft - (c + 1 + i)];              // p[N-n]
      if (c + 4 == half) y[3] = 0.f;                       // n == N/2 pairs with itself: e = p, o = 0
      const float xs[4] = {x.x, x.y, x.z, x.w};
      float eh[4], el[4], oh[4], ol[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        split2(xs[i] + y[i], eh[i], el[i]);
        split2(xs[i] - y[i], oh[i], ol[i]);
      }
      if (c + 4 == half) { oh[3] = 0.f; ol[3] = 0.f; }
      e_hi[q] = make_float4(eh[0], eh[1], eh[2], eh[3]);
      e_lo[q] = make_float4(el[0], el[1], el[2], el[3]);
      o_hi[q] = make_float4(oh[0], oh[1], oh[2], oh[3]);
      o_lo[q] = make_float4(ol[0], ol[1], ol[2], ol[3]);
    }
    if (p0 && threadIdx.x == 0) p0[f] = frame[0];
  }
}

// ------------------------------------------------------------------ K0h
// fp16 flavour of K0f for the 3xFP16 contraction.  fp16 carries tf32's 11 significant bits but only a 5-bit
// exponent, so each frame row is block-scaled: s = 14 - floor(log2(max(|e|,|o|))) puts the row maximum in
// [2^14, 2^15) (fp16 overflows at 65504), hi = fp16(v 2^s), lo = fp16(v 2^s - hi).  lo keeps its full 11 bits while
// |v 2^s| >= 2^-3, i.e. over 18 binades below the row maximum; under that it rounds on the fp16 subnormal grid,
// an absolute 2^-25 (2^-39 of the row maximum).  row_scale_inv[frame] = 2^-s is applied by the GEMM epilogue.
//
// Block = kFoldWarps consecutive frames of one segment.  Their n_fft + (kFoldWarps-1) hop RAW samples (float or PCM16)
// are staged in shared memory by ONE bulk async copy (cp.async.bulk + mbarrier: the TMA engine moves the 11-22 KB, no
// thread holds staging registers); blocks are persistent and double-buffered, so the copy of the next group is in
// flight while the warps fold this one.  Then one WARP owns one frame: fold, warp-shuffle max, scale, split, 8-byte
// stores -- no block barrier between staging and use, one per group to recycle the buffer.  Groups that touch the
// reflect padding (or are not 16-byte aligned) are filled by the threads instead.
// smem index = sample index: the mirrored run p[N-c-4..N-c-1] is one aligned vector load, the forward run
// p[c+1..c+4] straddles two.
constexpr int kFoldWarps = 8;

__device__ __forceinline__ uint32_t fs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool fs_mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}

template <typename TIn, bool kRaw = false>
__device__ __forceinline__ void fold4(const TIn* __restrict__ fr /* fr[i] = p[i] */, float gain, int n_fft, int half, int c,
                                      float (&e)[4], float (&o)[4]) {
  if constexpr (kRaw) gain = 1.f;                                                // integer PCM units: the multiply folds away
  float xs[4], ys[4];
  if constexpr (sizeof(TIn) == 4) {
    const float4 q0 = *reinterpret_cast<const float4*>(fr + c);                  // p[c .. c+3]
    const float4 q1 = *reinterpret_cast<const float4*>(fr + c + 4);              // p[c+4 .. c+7]
    const float4 m = *reinterpret_cast<const float4*>(fr + n_fft - c - 4);       // p[N-c-4 .. N-c-1]
    xs[0] = q0.y; xs[1] = q0.z; xs[2] = q0.w; xs[3] = q1.x;                      // p[n], n = c+1 .. c+4
    ys[0] = m.w; ys[1] = m.z; ys[2] = m.y; ys[3] = m.x;                          // p[N-n]
  } else {
    const short4 q0 = *reinterpret_cast<const short4*>(fr + c);
    const short4 q1 = *reinterpret_cast<const short4*>(fr + c + 4);
    const short4 m = *reinterpret_cast<const short4*>(fr + n_fft - c - 4);
    xs[0] = (float)q0.y * gain; xs[1] = (float)q0.z * gain; xs[2] = (float)q0.w * gain; xs[3] = (float)q1.x * gain;
    ys[0] = (float)m.w * gain; ys[1] = (float)m.z * gain; ys[2] = (float)m.y * gain; ys[3] = (float)m.x * gain;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    e[i] = xs[i] + ys[i];
    o[i] = xs[i] - ys[i];
  }
  if (c + 4 == half) { e[3] = xs[3]; o[3] = 0.f; }                               // n == N/2 pairs with itself
}

// TIn = float, or int16_t for PCM16 audio as the dataset stores it (sample = pcm * gain, gain = 1/32768:
// model/dataset.py:62 `audio.float().div_(32768.0)`; both steps are exact in fp32).
// kPerm (planes of the TWICE-folded contraction, rvb_stft_mel_folded2_f16): the columns of a row are ordered by the
// parity of n = c + 1 -- even n first (n = 2, 4, .., N/2 -> columns 0 .. N/4-1), then odd n (columns N/4 .. N/2-1).  A
// lane's four consecutive n are two even and two odd ones: two 4-byte stores per plane instead of one 8-byte store.
// kFixed (PCM16 with a power-of-two gain): no max pass.  e = x +- y of two int16 samples is an integer of at most 17
// bits; e / 4 lies inside the fp16 range and hi = fp16(e/4), lo = fp16(e/4 - hi) hold it EXACTLY (11 + 6 bits, the
// residual never below 2^-2), whatever the level of the frame -- so every row takes the same scale 4 * gain.  The
// planes differ from the block-scaled ones by an exact power of two per row, the contraction's result not at all.
template <typename TIn, bool kPerm, bool kFixed = false>
__global__ void __launch_bounds__(kFoldWarps * 32)
fold_split_f16_kernel(const TIn* __restrict__ audio, int64_t audio_ld, float gain, int n_seg, int n_samples, int pad,
                      int mode, int n_fft, int hop, int n_frames, int groups_per_seg, __half* __restrict__ a_hi,
                      __half* __restrict__ a_lo, float* __restrict__ row_scale_inv, float* __restrict__ p0) {
  extern __shared__ __align__(128) uint8_t fs_raw[];
  __shared__ __align__(8) uint64_t fs_bar[2];
  const int half = n_fft >> 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n_rows = (int64_t)n_seg * n_frames;
  const int64_t padded = (mode == RVB_PAD_NONE) ? n_samples : (int64_t)n_samples + 2 * pad;
  const int off = (mode == RVB_PAD_NONE) ? 0 : pad;
  const int span = n_fft + (kFoldWarps - 1) * hop;
  const int buf_elems = (span + 8 + 31) & ~31;             // + slack for the forward straddle of the last quad
  const int n_groups = n_seg * groups_per_seg;
  TIn* const buf0 = reinterpret_cast<TIn*>(fs_raw);
  const uint32_t bar0 = fs_smem_u32(&fs_bar[0]);           // buffer / barrier `which` = base + which * stride
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // where a group's samples start, and whether one aligned bulk copy can fetch them
  auto locate = [&](int grp, const TIn*& src) -> bool {
    const int b = grp / groups_per_seg;
    const int t0 = (grp - b * groups_per_seg) * kFoldWarps;
    const TIn* a = audio + (int64_t)b * audio_ld;
    const int64_t j0 = (int64_t)t0 * hop - off;
    src = a + j0;
    return j0 >= 0 && j0 + span <= n_samples && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) &&
           (((size_t)span * sizeof(TIn)) & 15u) == 0;
  };
  auto issue = [&](int grp, int which) {                     // thread 0 only
    const TIn* src;
    if (locate(grp, src)) {
      const uint32_t bytes = (uint32_t)((size_t)span * sizeof(TIn));
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * which), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(fs_smem_u32(buf0 + which * buf_elems)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes),
                     "r"(bar0 + 8 * which)
                   : "memory");
    }
  };

  uint32_t phases = 0u;                                    // bit `which` = parity the next wait on that barrier expects
  if (threadIdx.x == 0 && (int)blockIdx.x < n_groups) issue(blockIdx.x, 0);
  int it = 0;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x, ++it) {
    const int cur = it & 1;
    const int nxt = grp + gridDim.x;
    if (threadIdx.x == 0 && nxt < n_groups) issue(nxt, cur ^ 1);   // its previous readers passed the barrier below
    const int b = grp / groups_per_seg;
    const int t0 = (grp - b * groups_per_seg) * kFoldWarps;
    const TIn* src;
    TIn* S = buf0 + cur * buf_elems;
    if (locate(grp, src)) {
      while (!fs_mbar_try_wait(bar0 + 8 * cur, (phases >> cur) & 1u)) {}
      phases ^= 1u << cur;
    } else {
      const TIn* a = audio + (int64_t)b * audio_ld;
      const int64_t start = (int64_t)t0 * hop;
      for (int i = threadIdx.x; i < span; i += blockDim.x) {
        const int64_t pi = start + i;
        TIn v = (TIn)0;
        if (pi < padded) {
          int64_t j = (mode == RVB_PAD_NONE) ? pi : pi - pad;
          bool zero = false;
          if (j < 0) { if (mode == RVB_PAD_REFLECT) j = -j; else zero = true; }
          else if (j >= n_samples) { if (mode == RVB_PAD_REFLECT) j = 2 * (int64_t)(n_samples - 1) - j; else zero = true; }
          if (!zero) v = __ldg(a + j);
        }
        S[i] = v;
      }
      __syncthreads();
    }
    const int t = t0 + warp;
    if (t < n_frames) {                                      // warp-uniform
      const TIn* fr = S + warp * hop;                        // fr[i] = p[i] of frame t
      // two passes over the staged samples (max, then scale + split): keeping the 64 folded values of a lane in
      // registers instead costs 95 registers, halves the occupancy and is slower
      float ev[4], ov[4];
      int s = 0;
      float sc = 0.25f;
      if constexpr (!kFixed) {
        float mx = 0.f;
        for (int c = lane << 2; c < half; c += 128) {
          fold4<TIn>(fr, gain, n_fft, half, c, ev, ov);
#pragma unroll
          for (int i = 0; i < 4; ++i) mx = fmaxf(mx, fmaxf(fabsf(ev[i]), fabsf(ov[i])));   // fmaxf drops NaN: s stays finite
        }
        mx = warp_max(mx);
        // floor(log2(mx)) from the exponent field (subnormal rows: treated as 2^-126); all-zero rows: s = 0
        if (mx > 0.f) {
          int ex = (int)((__float_as_uint(mx) >> 23) & 0xff) - 127;
          ex = max(-126, min(ex, 127));
          s = max(-126, min(14 - ex, 126));
        }
        sc = __uint_as_float((unsigned)(s + 127) << 23);                    // 2^s, exact
      }
      const int64_t f = (int64_t)b * n_frames + t;
      uint2* e_hi = reinterpret_cast<uint2*>(a_hi + f * half);
      uint2* e_lo = reinterpret_cast<uint2*>(a_lo + f * half);
      uint2* o_hi = reinterpret_cast<uint2*>(a_hi + (n_rows + f) * half);
      uint2* o_lo = reinterpret_cast<uint2*>(a_lo + (n_rows + f) * half);
      // hi = fp16(v 2^s), lo = fp16(v 2^s - hi), two values per cvt.rn.f16x2
      auto split4 = [sc](const float (&v)[4], uint2& hi, uint2& lo) {
        const float a0 = v[0] * sc, a1 = v[1] * sc, a2 = v[2] * sc, a3 = v[3] * sc;
        const __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(a0 - f01.x, a1 - f01.y), l23 = __floats2half2_rn(a2 - f23.x, a3 - f23.y);
        hi.x = *reinterpret_cast<const uint32_t*>(&h01); hi.y = *reinterpret_cast<const uint32_t*>(&h23);
        lo.x = *reinterpret_cast<const uint32_t*>(&l01); lo.y = *reinterpret_cast<const uint32_t*>(&l23);
      };
      for (int c = lane << 2; c < half; c += 128) {
        fold4<TIn, kFixed>(fr, gain, n_fft, half, c, ev, ov);
        uint2 h, l;
        if constexpr (kPerm) {
          // n = c+1 .. c+4: (ev[1], ev[3]) are the even n -> pair c/2 of the first half of the row, (ev[0], ev[2]) the
          // odd n -> pair c/2 of the second half
          const int pe = c >> 2, po = (half >> 2) + (c >> 2);                 // in units of two halves (4 bytes)
          const float e2[4] = {ev[1], ev[3], ev[0], ev[2]}, o2[4] = {ov[1], ov[3], ov[0], ov[2]};
          split4(e2, h, l);
          reinterpret_cast<uint32_t*>(e_hi)[pe] = h.x; reinterpret_cast<uint32_t*>(e_hi)[po] = h.y;
          reinterpret_cast<uint32_t*>(e_lo)[pe] = l.x; reinterpret_cast<uint32_t*>(e_lo)[po] = l.y;
          split4(o2, h, l);
          reinterpret_cast<uint32_t*>(o_hi)[pe] = h.x; reinterpret_cast<uint32_t*>(o_hi)[po] = h.y;
          reinterpret_cast<uint32_t*>(o_lo)[pe] = l.x; reinterpret_cast<uint32_t*>(o_lo)[po] = l.y;
        } else {
          split4(ev, h, l); e_hi[c >> 2] = h; e_lo[c >> 2] = l;
          split4(ov, h, l); o_hi[c >> 2] = h; o_lo[c >> 2] = l;
        }
      }
      if (lane == 0) {
        row_scale_inv[f] = kFixed ? 4.f * gain : __uint_as_float((unsigned)(127 - s) << 23);      // 2^-s
        if (p0) p0[f] = (float)fr[0] * (sizeof(TIn) == 4 ? 1.f : gain);
      }
    }
    __syncthreads();                                         // everyone is done with bufs[cur]: it may be refilled
  }
}

// ------------------------------------------------------------------ K1b
// One warp per frame: plain fp32 dot products of the frame with one basis row pair.
__global__ void __launch_bounds__(256)
stft_bin_kernel(const float* __restrict__ sig_hi, const float* __restrict__ sig_lo, int n_seg, int rows_per_seg,
                int hop, int n_frames, const float* __restrict__ wcos_row, const float* __restrict__ wsin_row,
                int n_fft, int bin, int epilogue, float power, float* __restrict__ out0, int n_out_bins) {
  const int lane = threadIdx.x & 31;
  const int64_t frame = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (frame >= (int64_t)n_seg * n_frames) return;
  const int b = (int)(frame / n_frames), t = (int)(frame % n_frames);
  const int64_t base = ((int64_t)b * rows_per_seg + t) * hop;   // frame t starts at plane sample t*hop
  float re = 0.f, im = 0.f;
  for (int n = lane; n < n_fft; n += 32) {
    float s = __ldg(sig_hi + base + n) + __ldg(sig_lo + base + n);
    re = fmaf(s, __ldg(wcos_row + n), re);
    im = fmaf(s, __ldg(wsin_row + n), im);
  }
  re = warp_sum(re);
  im = warp_sum(im);
  if (lane == 0) stft_store(epilogue, power, re, im, out0, b, bin, t, n_out_bins, n_frames);
}

// ------------------------------------------------------------------ K2
// Input: power[frame][bin] time-major (the GEMM epilogue writes it that way with 16-byte stores).
// Block = kMelFR consecutive frames of one segment: their rows (kMelFR x n_bins floats, contiguous in HBM)
// are staged in smem with coalesced float4 loads; then thread <-> Mel band: it holds its band's weights in
// registers and walks its contiguous bin support in the smem rows (loop bound = longest support in the
// warp, so low-frequency warps finish in a few steps).  Output is coalesced in the time-major layout
// (thread <-> band) and sector-complete in the bin-major one.  log and the per-segment min/max keys are fused.
constexpr int kMelFR = 8;
constexpr int kMelThreads = 256;

__global__ void __launch_bounds__(kMelThreads)
mel_project_kernel(const float* __restrict__ power, int n_frames, int n_bins, const int32_t* __restrict__ band_lo,
                   const int32_t* __restrict__ band_len, const float* __restrict__ band_w, int max_len, int n_mels,
                   float log_offset, int layout, float* __restrict__ out, uint32_t* __restrict__ minmax) {
  extern __shared__ __align__(16) float rows[];            // [kMelFR][n_bins]
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kMelFR;
  const int nfr = min(kMelFR, n_frames - t0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* src = power + ((int64_t)b * n_frames + t0) * n_bins;
  const int n_val = nfr * n_bins;
  if ((n_bins & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(rows);
    for (int i = threadIdx.x; i < (n_val >> 2); i += kMelThreads) d4[i] = __ldg(s4 + i);
  } else {
    for (int i = threadIdx.x; i < n_val; i += kMelThreads) rows[i] = __ldg(src + i);
  }
  __syncthreads();

  float vmax = -INFINITY, vmin = INFINITY;
  bool seen_nan = false;
  for (int m0 = 0; m0 < n_mels; m0 += kMelThreads) {
    const int m = m0 + threadIdx.x;
    const bool m_ok = m < n_mels;
    const int lo = m_ok ? __ldg(band_lo + m) : 0;
    const int len = m_ok ? __ldg(band_len + m) : 0;
    int wlen = len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wlen = max(wlen, __shfl_xor_sync(kFull, wlen, o));
    float acc[kMelFR];
#pragma unroll
    for (int fr = 0; fr < kMelFR; ++fr) acc[fr] = 0.f;
    const float* base = rows + lo;
    const float* wcol = band_w + (m_ok ? m : 0);           // weights are stored [j][band]: coalesced, L1-resident
    for (int j = 0; j < wlen; ++j) {                       // warp-uniform bound = longest support in the warp
      const bool in = j < len;
      const float w = in ? __ldg(wcol + (int64_t)j * n_mels) : 0.f;
      const int jj = in ? j : 0;                           // stay inside the row; the weight is 0 there
#pragma unroll
      for (int fr = 0; fr < kMelFR; ++fr) acc[fr] = fmaf(w, base[fr * n_bins + jj], acc[fr]);
    }
    if (m_ok) {
#pragma unroll
      for (int fr = 0; fr < kMelFR; ++fr) {
        if (fr < nfr) {
          float v = acc[fr];
          if (log_offset >= 0.f) v = logf(v + log_offset);
          acc[fr] = v;
          vmax = fmaxf(vmax, v); vmin = fminf(vmin, v); seen_nan |= isnan(v);
        }
      }
      if (layout == RVB_LAYOUT_TIME_MAJOR) {
        float* dst = out + ((int64_t)b * n_frames + t0) * n_mels + m;
#pragma unroll
        for (int fr = 0; fr < kMelFR; ++fr)
          if (fr < nfr) dst[(int64_t)fr * n_mels] = acc[fr];
      } else {
        float* dst = out + ((int64_t)b * n_mels + m) * n_frames + t0;
        if (nfr == kMelFR && (n_frames & 3) == 0) {
          reinterpret_cast<float4*>(dst)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
          reinterpret_cast<float4*>(dst)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        } else {
#pragma unroll
          for (int fr = 0; fr < kMelFR; ++fr)
            if (fr < nfr) dst[fr] = acc[fr];
        }
      }
    }
  }
  if (minmax) {
    // torch.max / torch.min propagate NaN: encode it as the largest key on both sides.
    unsigned kmax = warp_max_u32(seen_nan ? 0xffffffffu : f2key(vmax));
    unsigned kmin = warp_max_u32(seen_nan ? 0xffffffffu : f2key(-vmin));
    __shared__ unsigned red[2][kMelThreads / 32];
    if (lane == 0) { red[0][warp] = kmin; red[1][warp] = kmax; }
    __syncthreads();
    if (threadIdx.x < 2) {
      unsigned k = 0;
#pragma unroll
      for (int w2 = 0; w2 < kMelThreads / 32; ++w2) k = max(k, red[threadIdx.x][w2]);
      atomicMax(minmax + 2 * b + threadIdx.x, k);
    }
  }
}

// ------------------------------------------------------------------ min/max of an arbitrary tensor
__global__ void __launch_bounds__(256)
minmax_kernel(const float* __restrict__ x, int64_t n_per_seg, uint32_t* __restrict__ minmax) {
  const int b = blockIdx.y;
  const float* p = x + (int64_t)b * n_per_seg;
  float vmax = -INFINITY, vmin = INFINITY;
  bool seen_nan = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_seg; i += (int64_t)gridDim.x * blockDim.x) {
    float v = __ldg(p + i);
    vmax = fmaxf(vmax, v); vmin = fminf(vmin, v); seen_nan |= isnan(v);
  }
  unsigned kmax = warp_max_u32(seen_nan ? 0xffffffffu : f2key(vmax));
  unsigned kmin = warp_max_u32(seen_nan ? 0xffffffffu : f2key(-vmin));
  __shared__ unsigned red[2][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = kmin; red[1][warp] = kmax; }
  __syncthreads();
  if (threadIdx.x < 2) {
    unsigned k = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) k = max(k, red[threadIdx.x][w]);
    atomicMax(minmax + 2 * b + threadIdx.x, k);
  }
}

// ------------------------------------------------------------------ K2m: after the fused Mel epilogue
// log via MUFU.LG2 (absolute error < 2^-21 outside [0.5, 2], 1 ulp inside -- three orders below the 1e-4 log-Mel
// tolerance): libm's logf is ~25 instructions and made both passes instruction-bound (profiles/r01d).  Both passes
// use the same function, so the normalised extrema stay exactly 0 and 1.
// __fmul_rn is never contracted into a following subtract, so both passes round identically.
__device__ __forceinline__ float fast_log(float x) { return __fmul_rn(__log2f(x), 0.693147182464599609375f); }

// mel[b][m][t] (bins-major, what MelSpectrogram.forward returns) -> per-segment min/max keys of log(mel + offset).
__global__ void __launch_bounds__(256)
logmel_minmax_kernel(const float* __restrict__ mel, int64_t n_per_seg, float log_offset, uint32_t* __restrict__ minmax) {
  const int b = blockIdx.y;
  const float* p = mel + (int64_t)b * n_per_seg;
  float vmax = -INFINITY, vmin = INFINITY;
  bool seen_nan = false;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((n_per_seg & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
    for (int64_t j = i0; j < (n_per_seg >> 2); j += stride) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(p) + j);
      const float v[4] = {fast_log(q.x + log_offset), fast_log(q.y + log_offset), fast_log(q.z + log_offset), fast_log(q.w + log_offset)};
#pragma unroll
      for (int e = 0; e < 4; ++e) { vmax = fmaxf(vmax, v[e]); vmin = fminf(vmin, v[e]); seen_nan |= isnan(v[e]); }
    }
  } else {
    for (int64_t j = i0; j < n_per_seg; j += stride) {
      const float v = fast_log(__ldg(p + j) + log_offset);
      vmax = fmaxf(vmax, v); vmin = fminf(vmin, v); seen_nan |= isnan(v);
    }
  }
  unsigned kmax = warp_max_u32(seen_nan ? 0xffffffffu : f2key(vmax));
  unsigned kmin = warp_max_u32(seen_nan ? 0xffffffffu : f2key(-vmin));
  __shared__ unsigned red[2][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = kmin; red[1][warp] = kmax; }
  __syncthreads();
  if (threadIdx.x < 2) {
    unsigned k = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) k = max(k, red[threadIdx.x][w]);
    atomicMax(minmax + 2 * b + threadIdx.x, k);
  }
}

__device__ __forceinline__ void decode_minmax(const uint32_t* __restrict__ minmax, int b, float& mn, float& mx);

// mel[b][m][t] -> out[b][t][m] = (log(mel + offset) - min) / (max - min)   (log only when minmax == nullptr):
// model/self_attention_VAT.py:1102, utils.py:100 and the .transpose(-1,-2) of :1104 in one pass.  Block = 32 frames x
// ALL bands: 128-byte row reads, and the 32 output rows form one contiguous n_mels*128-byte run (float4 stores).
constexpr int kTrFrames = 32;

__global__ void __launch_bounds__(256)
logmel_transpose_kernel(const float* __restrict__ mel, int n_mels, int n_frames, float log_offset,
                        const uint32_t* __restrict__ minmax, float* __restrict__ out) {
  extern __shared__ __align__(16) float tile[];            // [kTrFrames][n_mels]: already in output order
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kTrFrames;
  const int nt = min(kTrFrames, n_frames - t0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float mn = 0.f, den = 1.f;
  if (minmax) {
    float mx;
    decode_minmax(minmax, b, mn, mx);
    den = mx - mn;                                         // (x_max - x_min), utils.py:100
  }
  const float rden = 1.f / den;
  const bool rcp_ok = rcp_usable(rden);                    // block-uniform: constant / NaN segments take the real division
  const float* src = mel + (int64_t)b * n_mels * n_frames + t0;
  // one warp per band row (32 consecutive frames = 128 bytes); four rows in flight per warp to cover DRAM latency
  for (int m4 = warp; m4 < n_mels; m4 += 32) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int m = m4 + 8 * u;
      v[u] = (m < n_mels && lane < nt) ? __ldg(src + (int64_t)m * n_frames + lane) : 1.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int m = m4 + 8 * u;
      float w = v[u];
      if (log_offset >= 0.f) w = fast_log(w + log_offset);
      if (minmax) w = rcp_ok ? div_by<true>(w - mn, den, rden) : (w - mn) / den;
      if (m < n_mels && lane < nt) tile[lane * n_mels + m] = w;   // bank = (lane * n_mels + m) % 32: conflict-free, odd n_mels
    }
  }
  __syncthreads();
  float* dst = out + ((int64_t)b * n_frames + t0) * n_mels;
  const int n = nt * n_mels;
  if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
    for (int i = threadIdx.x; i < (n >> 2); i += 256)
      reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(tile)[i];
    for (int i = (n & ~3) + threadIdx.x; i < n; i += 256) dst[i] = tile[i];
  } else {
    for (int i = threadIdx.x; i < n; i += 256) dst[i] = tile[i];
  }
}

// ------------------------------------------------------------------ K2m + K3m in ONE pass: a cluster per segment
// The imagewise min/max (model/utils.py:96-97) is the only step of the front-end that couples a whole segment, and it
// is why the two kernels above read the Mel spectrogram twice.  Here a thread-block CLUSTER owns one segment: CTA r
// keeps frames [r*fpc, (r+1)*fpc) of log(mel + offset) in shared memory, already transposed into output order
// (229 x 80 floats = 73 KB per CTA for a 640-frame segment on 8 CTAs), the CTAs exchange their min/max keys through
// distributed shared memory (one remote store per peer + one cluster barrier), and every CTA then normalises its slab
// on the way out.  mel is read once, out is written once: 2 x N4 bytes per segment instead of 3 x N4, one launch
// instead of two (+ a memset).  Same fast_log, same keys, same (v - min) / (max - min) as the two-pass kernels:
// bit-identical results.
constexpr int kNormCluster = 8;              // portable cluster size
constexpr int kNormThreadsMax = 512;
constexpr int kNormItems = 4;                // vector path: (8 rows x 4 float4) items per warp and trip
constexpr int kNormRows = 4;                 // scalar path: band rows per warp and trip ...
constexpr int kNormChunks = 4;               // ... times 32-frame chunks: a CTA holds at most 128 frames

template <int kNormThreads>
__global__ void __launch_bounds__(kNormThreads)
logmel_normalise_cluster_kernel(const float* __restrict__ mel, const float* __restrict__ mel_b, int n_mels, int n_frames,
                                int frames_per_cta, int pitch, float log_offset, uint32_t* __restrict__ minmax_out,
                                float* __restrict__ out) {
  extern __shared__ __align__(16) float slab[];            // [frames_per_cta][pitch], pitch odd: conflict-free
  __shared__ unsigned red[2][kNormThreads / 32];
  __shared__ unsigned peer_keys[2][kNormCluster];          // [min|max][rank], written by the peers (DSMEM)
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  // A CTA may only be written through distributed shared memory once it has started executing: every CTA arrives
  // here, and waits for the others right before its remote stores (a split barrier: no stall in practice).
  cluster.barrier_arrive();
  const unsigned rank = cluster.block_rank();
  const int b = blockIdx.x / kNormCluster;
  const int t0 = (int)rank * frames_per_cta;
  const int nt = max(0, min(frames_per_cta, n_frames - t0));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kNormThreads / 32;

  // mel_b (optional): second plane of the twice-folded contraction (sin^2 part); the Mel spectrogram is mel + mel_b
  const int64_t seg_off = (int64_t)b * n_mels * n_frames + t0;
  const float* src = mel + seg_off;
  float vmax = -INFINITY, vmin = INFINITY;
  bool seen_nan = false;
  auto take = [&](float x, int t, int m) {                 // one value: log, extrema, transposed store
    const float w = fast_log(x + log_offset);
    vmax = fmaxf(vmax, w); vmin = fminf(vmin, w); seen_nan |= isnan(w);
    slab[t * pitch + m] = w;
  };
  const bool vec = (n_frames & 3) == 0 && (reinterpret_cast<uintptr_t>(mel) & 15u) == 0 &&
                   (mel_b == nullptr || (reinterpret_cast<uintptr_t>(mel_b) & 15u) == 0);
  if (vec) {
    // pass 1, 16-byte loads: a warp item is 8 band rows x 4 consecutive float4 (64 contiguous bytes per row: whole
    // sectors); lane = (row, quad).  kNormItems items per trip: up to 8 independent LDG.128 in flight per lane.
    const int nq = nt >> 2;                                // float4 per row (t0 and n_frames are multiples of 4)
    const int q_groups = (nq + 3) >> 2, r_groups = (n_mels + 7) >> 3;
    const int n_items = q_groups * r_groups;
    const int lr = lane >> 2, lq = lane & 3;
    for (int i0 = warp; i0 < n_items; i0 += kWarps * kNormItems) {
      float4 va[kNormItems], vb[kNormItems];
      int mm[kNormItems], qq[kNormItems];
#pragma unroll
      for (int u = 0; u < kNormItems; ++u) {
        const int i = i0 + u * kWarps;
        const int rg = i / q_groups, qg = i - rg * q_groups;
        mm[u] = rg * 8 + lr;
        qq[u] = qg * 4 + lq;
        const bool ok = i < n_items && mm[u] < n_mels && qq[u] < nq;
        if (!ok) mm[u] = -1;
        const int64_t o = (int64_t)mm[u] * n_frames + 4 * qq[u];
        va[u] = ok ? __ldg(reinterpret_cast<const float4*>(src + o)) : make_float4(1.f, 1.f, 1.f, 1.f);
        vb[u] = (ok && mel_b) ? __ldg(reinterpret_cast<const float4*>(mel_b + seg_off + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < kNormItems; ++u) {
        if (mm[u] >= 0) {
          float4 x = va[u];
          if (mel_b) { x.x = __fadd_rn(x.x, vb[u].x); x.y = __fadd_rn(x.y, vb[u].y); x.z = __fadd_rn(x.z, vb[u].z); x.w = __fadd_rn(x.w, vb[u].w); }
          const int t = 4 * qq[u];
          take(x.x, t, mm[u]); take(x.y, t + 1, mm[u]); take(x.z, t + 2, mm[u]); take(x.w, t + 3, mm[u]);
        }
      }
    }
  } else {
    // pass 1, any frame count: one warp-wide load = 32 consecutive frames of one band (128 bytes); a warp takes
    // kNormRows band rows per trip and all of their (up to kNormChunks) 32-frame chunks
    for (int m0 = warp; m0 < n_mels; m0 += kWarps * kNormRows) {
      float v[kNormRows][kNormChunks];
#pragma unroll
      for (int u = 0; u < kNormRows; ++u) {
        const int m = m0 + u * kWarps;
        const float* row = src + (int64_t)m * n_frames + lane;
        const float* row_b = mel_b ? mel_b + seg_off + (int64_t)m * n_frames + lane : nullptr;
#pragma unroll
        for (int c = 0; c < kNormChunks; ++c) {
          const bool ok = m < n_mels && c * 32 + lane < nt;
          v[u][c] = ok ? __ldg(row + c * 32) : 1.f;
          if (ok && row_b) v[u][c] = __fadd_rn(v[u][c], __ldg(row_b + c * 32));
        }
      }
#pragma unroll
      for (int u = 0; u < kNormRows; ++u) {
        const int m = m0 + u * kWarps;
#pragma unroll
        for (int c = 0; c < kNormChunks; ++c)
          if (m < n_mels && c * 32 + lane < nt) take(v[u][c], c * 32 + lane, m);
      }
    }
  }
  // block -> cluster reduction of the order-preserving keys (NaN = largest key on both sides, as torch.max / min)
  unsigned kmax = warp_max_u32(seen_nan ? 0xffffffffu : f2key(vmax));
  unsigned kmin = warp_max_u32(seen_nan ? 0xffffffffu : f2key(-vmin));
  if (lane == 0) { red[0][warp] = kmin; red[1][warp] = kmax; }
  __syncthreads();
  cluster.barrier_wait();                                  // completes the arrive at the top: all peers are running
  if (threadIdx.x < kNormCluster) {                        // thread r hands this CTA's keys to peer r
    unsigned k0 = 0, k1 = 0;
#pragma unroll
    for (int w2 = 0; w2 < kWarps; ++w2) { k0 = max(k0, red[0][w2]); k1 = max(k1, red[1][w2]); }
    unsigned* peer = cluster.map_shared_rank(&peer_keys[0][0], threadIdx.x);
    peer[rank] = k0;
    peer[kNormCluster + rank] = k1;
  }
  cluster.sync();                                          // release / acquire: the remote stores are visible
  unsigned gmin = 0, gmax = 0;
#pragma unroll
  for (int r = 0; r < kNormCluster; ++r) { gmin = max(gmin, peer_keys[0][r]); gmax = max(gmax, peer_keys[1][r]); }
  if (minmax_out && rank == 0 && threadIdx.x == 0) { minmax_out[2 * b] = gmin; minmax_out[2 * b + 1] = gmax; }
  float mn, mx;
  if (gmin == 0xffffffffu || gmax == 0xffffffffu) mn = mx = __int_as_float(0x7fc00000);
  else { mn = -key2f(gmin); mx = key2f(gmax); }
  const float den = mx - mn;                               // (x_max - x_min), utils.py:100
  const float rden = 1.f / den;

  // pass 2: the slab is one contiguous run of the output
  float* dst = out + ((int64_t)b * n_frames + t0) * n_mels;
  const int n = nt * n_mels;
  auto write = [&](auto fast) {
    constexpr bool kFast = decltype(fast)::value;
    if (pitch == n_mels && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
      for (int i = threadIdx.x; i < (n >> 2); i += kNormThreads) {
        float4 q = reinterpret_cast<const float4*>(slab)[i];
        q.x = div_by<kFast>(q.x - mn, den, rden); q.y = div_by<kFast>(q.y - mn, den, rden);
        q.z = div_by<kFast>(q.z - mn, den, rden); q.w = div_by<kFast>(q.w - mn, den, rden);
        reinterpret_cast<float4*>(dst)[i] = q;
      }
      for (int i = (n & ~3) + threadIdx.x; i < n; i += kNormThreads) dst[i] = div_by<kFast>(slab[i] - mn, den, rden);
    } else {
      for (int i = threadIdx.x; i < n; i += kNormThreads) {
        const int t = i / n_mels, m = i - t * n_mels;
        dst[i] = div_by<kFast>(slab[t * pitch + m] - mn, den, rden);
      }
    }
  };
  if (rcp_usable(rden)) write(std::true_type{}); else write(std::false_type{});   // constant / NaN segment: real division
}

// ------------------------------------------------------------------ K3
__device__ __forceinline__ void decode_minmax(const uint32_t* __restrict__ minmax, int b, float& mn, float& mx) {
  const unsigned kmin = __ldg(minmax + 2 * b), kmax = __ldg(minmax + 2 * b + 1);
  if (kmin == 0xffffffffu || kmax == 0xffffffffu) {
    mn = mx = __int_as_float(0x7fc00000);
  } else {
    mn = -key2f(kmin);
    mx = key2f(kmax);
  }
}

__global__ void __launch_bounds__(256)
normalise_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n_per_seg,
                 const uint32_t* __restrict__ minmax, int vec_ok) {
  const int b = blockIdx.y;
  float mn, mx;
  decode_minmax(minmax, b, mn, mx);
  const float den = mx - mn;                              // (x_max - x_min), utils.py:100
  const float* p = x + (int64_t)b * n_per_seg;
  float* q = y + (int64_t)b * n_per_seg;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec_ok) {
    const int64_t n4 = n_per_seg >> 2;
    for (int64_t j = i; j < n4; j += stride) {
      float4 v = __ldg(reinterpret_cast<const float4*>(p) + j);
      v.x = (v.x - mn) / den; v.y = (v.y - mn) / den; v.z = (v.z - mn) / den; v.w = (v.w - mn) / den;
      reinterpret_cast<float4*>(q)[j] = v;
    }
    for (int64_t j = (n4 << 2) + i; j < n_per_seg; j += stride) q[j] = (__ldg(p + j) - mn) / den;
  } else {
    for (int64_t j = i; j < n_per_seg; j += stride) q[j] = (__ldg(p + j) - mn) / den;
  }
}

// ------------------------------------------------------------------ K3f: Normalization('framewise')
// model/utils.py:85-92: per (segment, frame) min / max over the bins axis of x[b][bin][frame], (x - min)/(max - min),
// NaN -> 0 (a constant frame gives 0/0).  Thread <-> frame (lanes are consecutive frames: 128-byte rows), two
// passes over the bins; torch.max / torch.min propagate NaN, and so does the comparison chain here.
__global__ void __launch_bounds__(256)
normalise_framewise_kernel(const float* __restrict__ x, float* __restrict__ y, int n_bins, int n_frames) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_frames) return;
  const float* p = x + (int64_t)b * n_bins * n_frames + t;
  float* q = y + (int64_t)b * n_bins * n_frames + t;
  float mx = -INFINITY, mn = INFINITY;
  bool nan = false;
  for (int m = 0; m < n_bins; ++m) {
    const float v = __ldg(p + (int64_t)m * n_frames);
    mx = fmaxf(mx, v); mn = fminf(mn, v); nan |= isnan(v);
  }
  if (nan) mx = mn = __int_as_float(0x7fc00000);
  const float den = mx - mn;
  for (int m = 0; m < n_bins; ++m) {
    const float o = (__ldg(p + (int64_t)m * n_frames) - mn) / den;
    q[(int64_t)m * n_frames] = isnan(o) ? 0.f : o;                      // output[torch.isnan(output)] = 0
  }
}

}  // namespace rvb

using namespace rvb;

extern "C" int rvb_pad_split(const float* audio, int64_t audio_ld, int n_seg, int n_samples, int pad, int pad_mode,
                             float* sig_hi, float* sig_lo, int rows_per_seg, int hop, rvb_stream_t stream) {
  RVB_REQUIRE(audio && sig_hi && sig_lo, "rvb_pad_split: null pointer");
  RVB_REQUIRE(n_seg > 0 && n_samples > 0 && hop > 0 && rows_per_seg > 0, "rvb_pad_split: bad shape");
  RVB_REQUIRE(pad_mode >= RVB_PAD_REFLECT && pad_mode <= RVB_PAD_NONE, "rvb_pad_split: bad pad_mode %d", pad_mode);
  RVB_REQUIRE(hop % 4 == 0, "rvb_pad_split: hop %d must be a multiple of 4", hop);
  RVB_REQUIRE((reinterpret_cast<uintptr_t>(sig_hi) & 15u) == 0 && (reinterpret_cast<uintptr_t>(sig_lo) & 15u) == 0,
              "rvb_pad_split: planes must be 16-byte aligned");
  if (pad_mode == RVB_PAD_REFLECT) {
    // model/Spectrogram.py:214-215 raises for n < pad; ReflectionPad1d itself needs pad < n.
    RVB_REQUIRE(n_samples > pad, "rvb_pad_split: reflect padding %d needs more than %d samples", pad, n_samples);
  }
  const int64_t padded = (pad_mode == RVB_PAD_NONE) ? n_samples : (int64_t)n_samples + 2 * pad;
  const int64_t plane = (int64_t)rows_per_seg * hop;
  RVB_REQUIRE(plane >= padded, "rvb_pad_split: rows_per_seg*hop = %lld < padded length %lld", (long long)plane,
              (long long)padded);
  int64_t bx = (plane / 4 + 255) / 256;
  if (bx > 1024) bx = 1024;
  dim3 grid((unsigned)bx, (unsigned)n_seg);
  pad_split_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(audio, audio_ld, n_samples, pad, pad_mode, sig_hi, sig_lo,
                                                           plane);
  count_launch();
  return check_launch("pad_split_kernel");
}

extern "C" int rvb_fold_split(const float* audio, int64_t audio_ld, int n_seg, int n_samples, int pad, int pad_mode,
                              int n_fft, int hop, int n_frames, float* a_hi, float* a_lo, float* p0,
                              rvb_stream_t stream) {
  RVB_REQUIRE(audio && a_hi && a_lo, "rvb_fold_split: null pointer");
  RVB_REQUIRE((reinterpret_cast<uintptr_t>(a_hi) & 15u) == 0 && (reinterpret_cast<uintptr_t>(a_lo) & 15u) == 0,
              "rvb_fold_split: planes must be 16-byte aligned");
  RVB_REQUIRE(n_seg > 0 && n_samples > 0 && hop > 0 && n_frames > 0, "rvb_fold_split: bad shape");
  RVB_REQUIRE(n_fft >= 64 && n_fft % 64 == 0 && n_fft <= 32768, "rvb_fold_split: n_fft %d must be a multiple of 64", n_fft);
  RVB_REQUIRE(pad_mode >= RVB_PAD_REFLECT && pad_mode <= RVB_PAD_NONE, "rvb_fold_split: bad pad_mode %d", pad_mode);
  if (pad_mode == RVB_PAD_REFLECT)
    RVB_REQUIRE(n_samples > pad, "rvb_fold_split: reflect padding %d needs more than %d samples", pad, n_samples);
  const int64_t padded = (pad_mode == RVB_PAD_NONE) ? n_samples : (int64_t)n_samples + 2 * pad;
  RVB_REQUIRE((int64_t)(n_frames - 1) * hop + n_fft <= padded, "rvb_fold_split: %d frames do not fit %lld samples",
              n_frames, (long long)padded);
  const int64_t n_rows = (int64_t)n_seg * n_frames;
  const size_t smem = (size_t)(n_fft + 4) * sizeof(float);
  if (smem > 48 * 1024)
    RVB_CUDA(cudaFuncSetAttribute(fold_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t cap = 148 * 8 * 4;
  const unsigned grid = (unsigned)(n_rows < cap ? n_rows : cap);
  fold_split_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(audio, audio_ld, n_seg, n_samples, pad, pad_mode, n_fft,
                                                               hop, n_frames, a_hi, a_lo, p0);
  count_launch();
  return check_launch("fold_split_kernel");
}

template <typename TIn, bool kPerm = false>
static int launch_fold_split_f16(const char* who, const TIn* audio, int64_t audio_ld, float gain, int n_seg,
                                 int n_samples, int pad, int pad_mode, int n_fft, int hop, int n_frames, void* a_hi,
                                 void* a_lo, float* row_scale_inv, float* p0, rvb_stream_t stream) {
  RVB_REQUIRE(audio && a_hi && a_lo && row_scale_inv, "%s: null pointer", who);
  RVB_REQUIRE((reinterpret_cast<uintptr_t>(a_hi) & 15u) == 0 && (reinterpret_cast<uintptr_t>(a_lo) & 15u) == 0,
              "%s: planes must be 16-byte aligned", who);
  RVB_REQUIRE(n_seg > 0 && n_samples > 0 && hop > 0 && n_frames > 0, "%s: bad shape", who);
  RVB_REQUIRE(n_fft >= 128 && n_fft % 128 == 0 && n_fft <= 32768, "%s: n_fft %d must be a multiple of 128", who, n_fft);
  RVB_REQUIRE(pad_mode >= RVB_PAD_REFLECT && pad_mode <= RVB_PAD_NONE, "%s: bad pad_mode %d", who, pad_mode);
  if (pad_mode == RVB_PAD_REFLECT)
    RVB_REQUIRE(n_samples > pad, "%s: reflect padding %d needs more than %d samples", who, pad, n_samples);
  const int64_t padded = (pad_mode == RVB_PAD_NONE) ? n_samples : (int64_t)n_samples + 2 * pad;
  RVB_REQUIRE((int64_t)(n_frames - 1) * hop + n_fft <= padded, "%s: %d frames do not fit %lld samples", who, n_frames,
              (long long)padded);
  RVB_REQUIRE(hop % 4 == 0, "%s: hop %d must be a multiple of 4", who, hop);
  const int64_t span = n_fft + (kFoldWarps - 1) * (int64_t)hop;
  const size_t smem = 2 * (size_t)((span + 8 + 31) & ~31) * sizeof(TIn);     // two raw staging buffers
  RVB_REQUIRE(smem <= 200 * 1024, "%s: n_fft %d with hop %d needs %zu bytes of shared memory", who, n_fft, hop, smem);
  RVB_REQUIRE(!kPerm || n_fft % 256 == 0, "%s: n_fft %d must be a multiple of 256", who, n_fft);
  // PCM16 with a power-of-two gain (1/32768): fixed-scale planes, no max pass
  int gexp = 0;
  const bool fixed = sizeof(TIn) == 2 && gain > 0.f && std::frexp(gain, &gexp) == 0.5f && gexp > -100 && gexp < 100 &&
                     !getenv("RVB_NO_FIXED_SCALE");
  auto kernel = fixed ? fold_split_f16_kernel<TIn, kPerm, sizeof(TIn) == 2> : fold_split_f16_kernel<TIn, kPerm, false>;
  if (smem > 48 * 1024)
    RVB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int groups_per_seg = (n_frames + kFoldWarps - 1) / kFoldWarps;
  const int64_t n_groups = (int64_t)n_seg * groups_per_seg;
  RVB_REQUIRE(n_groups < (1ll << 31), "%s: too many frames", who);
  // persistent blocks: as many as stay resident (8 x 256 threads or the shared memory, whichever binds)
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  const int64_t cap = (int64_t)148 * per_sm;
  const unsigned grid = (unsigned)(n_groups < cap ? n_groups : cap);
  kernel<<<grid, kFoldWarps * 32, smem, (cudaStream_t)stream>>>(
      audio, audio_ld, gain, n_seg, n_samples, pad, pad_mode, n_fft, hop, n_frames, groups_per_seg,
      static_cast<__half*>(a_hi), static_cast<__half*>(a_lo), row_scale_inv, p0);
  count_launch();
  return check_launch("fold_split_f16_kernel");
}

extern "C" int rvb_fold_split_f16(const float* audio, int64_t audio_ld, int n_seg, int n_samples, int pad,
                                  int pad_mode, int n_fft, int hop, int n_frames, void* a_hi, void* a_lo,
                                  float* row_scale_inv, float* p0, rvb_stream_t stream) {
  return launch_fold_split_f16<float>("rvb_fold_split_f16", audio, audio_ld, 1.f, n_seg, n_samples, pad, pad_mode,
                                      n_fft, hop, n_frames, a_hi, a_lo, row_scale_inv, p0, stream);
}

extern "C" int rvb_fold_split_f16_pcm16(const int16_t* audio, int64_t audio_ld, float gain, int n_seg, int n_samples,
                                        int pad, int pad_mode, int n_fft, int hop, int n_frames, void* a_hi,
                                        void* a_lo, float* row_scale_inv, float* p0, rvb_stream_t stream) {
  return launch_fold_split_f16<int16_t>("rvb_fold_split_f16_pcm16", audio, audio_ld, gain, n_seg, n_samples, pad,
                                        pad_mode, n_fft, hop, n_frames, a_hi, a_lo, row_scale_inv, p0, stream);
}

extern "C" int rvb_fold_split2_f16(const float* audio, int64_t audio_ld, int n_seg, int n_samples, int pad,
                                   int pad_mode, int n_fft, int hop, int n_frames, void* a_hi, void* a_lo,
                                   float* row_scale_inv, rvb_stream_t stream) {
  return launch_fold_split_f16<float, true>("rvb_fold_split2_f16", audio, audio_ld, 1.f, n_seg, n_samples, pad, pad_mode,
                                            n_fft, hop, n_frames, a_hi, a_lo, row_scale_inv, nullptr, stream);
}

extern "C" int rvb_fold_split2_f16_pcm16(const int16_t* audio, int64_t audio_ld, float gain, int n_seg, int n_samples,
                                         int pad, int pad_mode, int n_fft, int hop, int n_frames, void* a_hi,
                                         void* a_lo, float* row_scale_inv, rvb_stream_t stream) {
  return launch_fold_split_f16<int16_t, true>("rvb_fold_split2_f16_pcm16", audio, audio_ld, gain, n_seg, n_samples, pad,
                                              pad_mode, n_fft, hop, n_frames, a_hi, a_lo, row_scale_inv, nullptr, stream);
}

extern "C" int rvb_stft_bin(const float* sig_hi, const float* sig_lo, int n_seg, int rows_per_seg, int hop,
                            int n_frames, const float* wcos_row, const float* wsin_row, int n_fft, int bin,
                            int epilogue, float power, float* out0, int n_out_bins, rvb_stream_t stream) {
  RVB_REQUIRE(sig_hi && sig_lo && wcos_row && wsin_row && out0, "rvb_stft_bin: null pointer");
  RVB_REQUIRE(n_seg > 0 && n_frames > 0 && bin >= 0 && bin < n_out_bins, "rvb_stft_bin: bad shape");
  RVB_REQUIRE(epilogue_ok(epilogue), "rvb_stft_bin: bad epilogue %d", epilogue);
  RVB_REQUIRE((int64_t)(n_frames - 1) * hop + n_fft <= (int64_t)rows_per_seg * hop, "rvb_stft_bin: plane too short");
  const int64_t frames = (int64_t)n_seg * n_frames;
  stft_bin_kernel<<<(unsigned)((frames + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      sig_hi, sig_lo, n_seg, rows_per_seg, hop, n_frames, wcos_row, wsin_row, n_fft, bin, epilogue, power, out0,
      n_out_bins);
  count_launch();
  return check_launch("stft_bin_kernel");
}

extern "C" int rvb_mel_project(const float* power, int n_seg, int n_frames, int n_bins, const int32_t* band_lo,
                               const int32_t* band_len, const float* band_w, int max_len, int n_mels, float log_offset,
                               int layout, float* out, uint32_t* minmax, rvb_stream_t stream) {
  RVB_REQUIRE(power && band_lo && band_len && band_w && out, "rvb_mel_project: null pointer");
  RVB_REQUIRE(n_seg > 0 && n_frames > 0 && n_mels > 0 && n_bins > 0, "rvb_mel_project: bad shape");
  RVB_REQUIRE(max_len > 0 && max_len <= n_bins, "rvb_mel_project: bad max_len %d", max_len);
  RVB_REQUIRE(layout == RVB_LAYOUT_BINS_MAJOR || layout == RVB_LAYOUT_TIME_MAJOR, "rvb_mel_project: bad layout");
  const size_t smem = (size_t)kMelFR * n_bins * sizeof(float);
  RVB_REQUIRE(smem <= 200 * 1024, "rvb_mel_project: n_bins %d too large", n_bins);
  if (minmax) RVB_CUDA(cudaMemsetAsync(minmax, 0, sizeof(uint32_t) * 2 * n_seg, (cudaStream_t)stream));
  if (smem > 48 * 1024)
    RVB_CUDA(cudaFuncSetAttribute(mel_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((n_frames + kMelFR - 1) / kMelFR), (unsigned)n_seg);
  mel_project_kernel<<<grid, kMelThreads, smem, (cudaStream_t)stream>>>(power, n_frames, n_bins, band_lo, band_len,
                                                                       band_w, max_len, n_mels, log_offset, layout, out,
                                                                       minmax);
  count_launch();
  return check_launch("mel_project_kernel");
}

extern "C" int rvb_logmel_minmax(const float* mel, int n_seg, int64_t n_per_seg, float log_offset, uint32_t* minmax,
                                 rvb_stream_t stream) {
  RVB_REQUIRE(mel && minmax, "rvb_logmel_minmax: null pointer");
  RVB_REQUIRE(n_seg > 0 && n_per_seg > 0 && log_offset >= 0.f, "rvb_logmel_minmax: bad argument");
  RVB_CUDA(cudaMemsetAsync(minmax, 0, sizeof(uint32_t) * 2 * n_seg, (cudaStream_t)stream));
  int64_t bx = (n_per_seg + 256 * 8 - 1) / (256 * 8);
  if (bx > 296) bx = 296;
  dim3 grid((unsigned)bx, (unsigned)n_seg);
  logmel_minmax_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mel, n_per_seg, log_offset, minmax);
  count_launch();
  return check_launch("logmel_minmax_kernel");
}

extern "C" int rvb_logmel_transpose(const float* mel, int n_seg, int n_mels, int n_frames, float log_offset,
                                    const uint32_t* minmax, float* out, rvb_stream_t stream) {
  RVB_REQUIRE(mel && out, "rvb_logmel_transpose: null pointer");
  RVB_REQUIRE(n_seg > 0 && n_mels > 0 && n_frames > 0 && n_seg <= 65535, "rvb_logmel_transpose: bad shape");
  const size_t smem = (size_t)kTrFrames * n_mels * sizeof(float);
  RVB_REQUIRE(smem <= 200 * 1024, "rvb_logmel_transpose: n_mels %d too large", n_mels);
  if (smem > 48 * 1024)
    RVB_CUDA(cudaFuncSetAttribute(logmel_transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((n_frames + kTrFrames - 1) / kTrFrames), (unsigned)n_seg);
  logmel_transpose_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(mel, n_mels, n_frames, log_offset, minmax, out);
  count_launch();
  return check_launch("logmel_transpose_kernel");
}

extern "C" int rvb_logmel_normalise(const float* mel, const float* mel_b, int n_seg, int n_mels, int n_frames,
                                    float log_offset, uint32_t* minmax, float* out, rvb_stream_t stream) {
  RVB_REQUIRE(mel && minmax && out, "rvb_logmel_normalise: null pointer");
  RVB_REQUIRE(n_seg > 0 && n_mels > 0 && n_frames > 0 && n_seg <= 65535 && log_offset >= 0.f,
              "rvb_logmel_normalise: bad argument");
  // frames per CTA: a multiple of 4 keeps every slab of the output 16-byte aligned
  const int fpc = (((n_frames + kNormCluster - 1) / kNormCluster) + 3) & ~3;
  const int pitch = n_mels | 1;
  const size_t smem = (size_t)fpc * pitch * sizeof(float);
  static const bool no_fusion = [] { const char* e = getenv("RVB_NO_NORM_FUSION"); return e && *e && *e != '0'; }();
  if (smem > 200 * 1024 || fpc > 32 * kNormChunks || no_fusion) {   // a segment that does not fit 8 SMs: two passes
    RVB_REQUIRE(!mel_b, "rvb_logmel_normalise: the two-pass path takes one plane (add the planes first)");
    int rc = rvb_logmel_minmax(mel, n_seg, (int64_t)n_mels * n_frames, log_offset, minmax, stream);
    if (rc != RVB_OK) return rc;
    return rvb_logmel_transpose(mel, n_seg, n_mels, n_frames, log_offset, minmax, out, stream);
  }
  // RVB_NORM_THREADS=256: 8 warps per CTA (16 K registers): a CTA then fits beside a resident contraction CTA of
  // another stream (with RVB_FOLD2_STAGES=3 leaving it 82 KB of shared memory)
  static const int n_threads = [] { const char* e = getenv("RVB_NORM_THREADS"); return (e && atoi(e) == 256) ? 256 : kNormThreadsMax; }();
  auto kernel = n_threads == 256 ? logmel_normalise_cluster_kernel<256> : logmel_normalise_cluster_kernel<kNormThreadsMax>;
  {
    static size_t smem_set[kMaxDevices] = {};              // the attribute call is not free: once per device and new maximum
    const int slot = device_slot();
    std::lock_guard<std::mutex> g(attr_mutex());
    if (smem > 48 * 1024 && smem > smem_set[slot]) {
      RVB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smem_set[slot] = smem;
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)n_seg * kNormCluster);
  cfg.blockDim = dim3(n_threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kNormCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RVB_CUDA(cudaLaunchKernelEx(&cfg, kernel, mel, mel_b, n_mels, n_frames, fpc, pitch, log_offset, minmax, out));
  count_launch();
  return check_launch("logmel_normalise_cluster_kernel");
}

extern "C" int rvb_minmax(const float* x, int n_seg, int64_t n_per_seg, uint32_t* minmax, rvb_stream_t stream) {
  RVB_REQUIRE(x && minmax, "rvb_minmax: null pointer");
  RVB_REQUIRE(n_seg > 0 && n_per_seg > 0, "rvb_minmax: bad shape");
  RVB_CUDA(cudaMemsetAsync(minmax, 0, sizeof(uint32_t) * 2 * n_seg, (cudaStream_t)stream));
  int64_t bx = (n_per_seg + 256 * 8 - 1) / (256 * 8);
  if (bx > 296) bx = 296;
  dim3 grid((unsigned)bx, (unsigned)n_seg);
  minmax_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n_per_seg, minmax);
  count_launch();
  return check_launch("minmax_kernel");
}

extern "C" int rvb_normalise(const float* x, float* y, int n_seg, int64_t n_per_seg, const uint32_t* minmax,
                             rvb_stream_t stream) {
  RVB_REQUIRE(x && y && minmax, "rvb_normalise: null pointer");
  RVB_REQUIRE(n_seg > 0 && n_per_seg > 0, "rvb_normalise: bad shape");
  const int vec = ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15u) == 0) &&
                  (n_per_seg % 4 == 0);
  int64_t bx = (n_per_seg + 256 * 8 - 1) / (256 * 8);
  if (bx > 296) bx = 296;
  dim3 grid((unsigned)bx, (unsigned)n_seg);
  normalise_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, n_per_seg, minmax, vec);
  count_launch();
  return check_launch("normalise_kernel");
}

extern "C" int rvb_normalise_framewise(const float* x, float* y, int n_seg, int n_bins, int n_frames,
                                       rvb_stream_t stream) {
  RVB_REQUIRE(x && y, "rvb_normalise_framewise: null pointer");
  RVB_REQUIRE(n_seg > 0 && n_bins > 0 && n_frames > 0 && n_seg <= 65535, "rvb_normalise_framewise: bad shape");
  dim3 grid((unsigned)((n_frames + 255) / 256), (unsigned)n_seg);
  normalise_framewise_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, n_bins, n_frames);
  count_launch();
  return check_launch("normalise_framewise_kernel");
}

// ------------------------------------------------------------------ D1: note decoding (caller side)
// model/decoding.py:4-55 extract_notes_wo_velocity: a note starts where the thresholded onset roll rises (and, rule1,
// the frame roll is on) and ends at the first later frame where neither roll is on.  The reference walks every note
// with a Python while loop and two .item() calls per frame.  Here one thread per pitch scans its column backwards
// (rows are contiguous over the pitches: coalesced), carrying "first inactive frame at or after t".
namespace rvb {
__global__ void __launch_bounds__(128)
note_offsets_kernel(const float* __restrict__ onsets, const float* __restrict__ frames, int T, int P, float onset_thr,
                    float frame_thr, int rule1, uint8_t* __restrict__ start, int32_t* __restrict__ offset) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  int next_inactive = T;
  bool on_t = __ldg(onsets + (int64_t)(T - 1) * P + p) > onset_thr;
  for (int t = T - 1; t >= 0; --t) {
    const bool fr_t = __ldg(frames + (int64_t)t * P + p) > frame_thr;
    const bool on_prev = (t > 0) ? (__ldg(onsets + (int64_t)(t - 1) * P + p) > onset_thr) : false;
    if (!(on_t || fr_t)) next_inactive = t;
    // uint8 arithmetic of the reference: onsets[t] - onsets[t-1] == 1  <=>  on now and off before (row 0: on now)
    const bool rises = on_t && !on_prev;
    start[(int64_t)t * P + p] = (rises && (!rule1 || fr_t)) ? 1 : 0;
    offset[(int64_t)t * P + p] = next_inactive;
    on_t = on_prev;
  }
}
}  // namespace rvb

extern "C" int rvb_note_offsets(const float* onsets, const float* frames, int n_frames, int n_pitches,
                                float onset_threshold, float frame_threshold, int rule1, uint8_t* start,
                                int32_t* offset, rvb_stream_t stream) {
  RVB_REQUIRE(onsets && frames && start && offset, "rvb_note_offsets: null pointer");
  RVB_REQUIRE(n_frames > 0 && n_pitches > 0, "rvb_note_offsets: bad shape");
  rvb::note_offsets_kernel<<<(unsigned)((n_pitches + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      onsets, frames, n_frames, n_pitches, onset_threshold, frame_threshold, rule1, start, offset);
  rvb::count_launch();
  return rvb::check_launch("note_offsets_kernel");
}

// ---------------------------------------------------------------- K0x: pad + parity split of the raw PCM16 signal
// The fused contraction (rvb_stft_mel_fused_pcm16, rvb_stft_gemm.cu) folds, scales and hi/lo-splits the frame rows
// itself.  What it reads is the reflect-padded signal (model/Spectrogram.py:209-218) split by SAMPLE PARITY -- the
// twice-folded contraction runs one chain over the even n and one over the odd n of a frame, and frames start at even
// samples -- and stored in OFFSET BINARY (u = x + 32768): the converter turns a sample into a float with one byte
// permute (0x4AC00000 | u is the float 1.5 * 2^22 + u / 2).  2 bytes per sample: the materialised fp16 frame planes of
// K0q are 16.  planes: [2][n_seg][plane_len] uint16, plane q element i = padded sample 2 i + q (32768 past the end).
namespace rvb {

__global__ void __launch_bounds__(256)
pad_parity_pcm16_kernel(const int16_t* __restrict__ audio, int64_t audio_ld, int n_seg, int n_samples, int pad, int mode,
                        uint16_t* __restrict__ planes, int64_t plane_len) {
  const int64_t groups_per_seg = plane_len >> 3;           // 8 plane elements of each parity = 16 padded samples
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= groups_per_seg * n_seg) return;
  const int b = (int)(g / groups_per_seg);
  const int64_t i0 = (g - (int64_t)b * groups_per_seg) << 3;
  const int64_t padded = (mode == RVB_PAD_NONE) ? n_samples : (int64_t)n_samples + 2 * pad;
  const int off = (mode == RVB_PAD_NONE) ? 0 : pad;
  const int16_t* a = audio + (int64_t)b * audio_ld;
  const int64_t j0 = 2 * i0 - off;                          // source index of the first of the 16 samples
  uint4 even, odd;
  if (j0 >= 0 && j0 + 16 <= n_samples && ((reinterpret_cast<uintptr_t>(a + j0) & 15u) == 0)) {
    const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(a + j0));
    const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(a + j0) + 1);
    // (s0 s1)(s2 s3) -> even (s0 s2), odd (s1 s3); x ^ 0x8000 = x + 32768 mod 2^16
    even.x = __byte_perm(v0.x, v0.y, 0x5410) ^ 0x80008000u; odd.x = __byte_perm(v0.x, v0.y, 0x7632) ^ 0x80008000u;
    even.y = __byte_perm(v0.z, v0.w, 0x5410) ^ 0x80008000u; odd.y = __byte_perm(v0.z, v0.w, 0x7632) ^ 0x80008000u;
    even.z = __byte_perm(v1.x, v1.y, 0x5410) ^ 0x80008000u; odd.z = __byte_perm(v1.x, v1.y, 0x7632) ^ 0x80008000u;
    even.w = __byte_perm(v1.z, v1.w, 0x5410) ^ 0x80008000u; odd.w = __byte_perm(v1.z, v1.w, 0x7632) ^ 0x80008000u;
  } else {
    uint32_t w[8];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int64_t pi = 2 * i0 + e;
      int v = 0;
      if (pi < padded) {
        int64_t j = pi - off;
        bool zero = false;
        if (j < 0) { if (mode == RVB_PAD_REFLECT) j = -j; else zero = true; }
        else if (j >= n_samples) { if (mode == RVB_PAD_REFLECT) j = 2 * (int64_t)(n_samples - 1) - j; else zero = true; }
        if (!zero && j >= 0 && j < n_samples) v = a[j];
      }
      const uint32_t u = (uint32_t)(v + 32768) & 0xffffu;
      // e = 2 m + q: element m of parity q; two elements per 32-bit word
      const int q = e & 1, m = e >> 1;
      uint32_t& word = w[q * 4 + (m >> 1)];
      word = (m & 1) ? (word | (u << 16)) : u;
    }
    even = make_uint4(w[0], w[1], w[2], w[3]);
    odd = make_uint4(w[4], w[5], w[6], w[7]);
  }
  uint16_t* pe = planes + (int64_t)b * plane_len + i0;
  uint16_t* po = planes + ((int64_t)n_seg + b) * plane_len + i0;
  *reinterpret_cast<uint4*>(pe) = even;
  *reinterpret_cast<uint4*>(po) = odd;
}

}  // namespace rvb

extern "C" int64_t rvb_parity_plane_len(int n_samples, int pad, int pad_mode, int n_fft, int hop, int n_frames) {
  const int64_t padded = (pad_mode == RVB_PAD_NONE) ? n_samples : (int64_t)n_samples + 2 * pad;
  int64_t need = (int64_t)(hop / 2) * (n_frames > 0 ? n_frames - 1 : 0) + n_fft / 2 + 8;   // + the converter's look-ahead
  if ((padded + 1) / 2 > need) need = (padded + 1) / 2;
  return (need + 7) & ~(int64_t)7;
}

extern "C" int rvb_pad_parity_pcm16(const int16_t* audio, int64_t audio_ld, int n_seg, int n_samples, int pad, int pad_mode,
                                    uint16_t* planes, int64_t plane_len, rvb_stream_t stream) {
  RVB_REQUIRE(audio && planes, "rvb_pad_parity_pcm16: null pointer");
  RVB_REQUIRE(n_seg > 0 && n_samples > 0 && pad >= 0 && plane_len > 0 && (plane_len & 7) == 0,
              "rvb_pad_parity_pcm16: bad shape (plane_len must be a positive multiple of 8)");
  RVB_REQUIRE(pad_mode == RVB_PAD_REFLECT || pad_mode == RVB_PAD_CONSTANT || pad_mode == RVB_PAD_NONE,
              "rvb_pad_parity_pcm16: bad pad mode %d", pad_mode);
  RVB_REQUIRE(pad_mode != RVB_PAD_REFLECT || pad < n_samples, "rvb_pad_parity_pcm16: reflect padding needs pad < n_samples");
  RVB_REQUIRE((reinterpret_cast<uintptr_t>(planes) & 15u) == 0, "rvb_pad_parity_pcm16: planes must be 16-byte aligned");
  const int64_t groups = (plane_len >> 3) * n_seg;
  RVB_REQUIRE(groups < (1ll << 31) * 256, "rvb_pad_parity_pcm16: too many samples");
  rvb::pad_parity_pcm16_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      audio, audio_ld, n_seg, n_samples, pad, pad_mode, planes, plane_len);
  rvb::count_launch();
  return rvb::check_launch("pad_parity_pcm16_kernel");
}
