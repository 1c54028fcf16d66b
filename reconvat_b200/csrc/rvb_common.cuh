// Shared device/host helpers for the reconvat_b200 kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <mutex>

#include "../../include/rvb.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "reconvat_b200 kernels are written for sm_100a (B200) only"
#endif

namespace rvb {

// ---- host-side error plumbing (thread-local last-error string, see rvb_last_error) ----
void set_error(const char* fmt, ...);
int check_launch(const char* what);           // cudaGetLastError -> RVB_ERR_LAUNCH

// Function attributes (opt-in dynamic shared memory, cluster occupancy) and the SM count belong to a DEVICE, not to the
// process: host-side caches of them are arrays indexed by device_slot() and are read / written under attr_mutex()
// (one process may drive several GPUs from several threads -- nn.DataParallel).
constexpr int kMaxDevices = 64;
int device_slot();                            // cudaGetDevice(), folded into [0, kMaxDevices)
std::mutex& attr_mutex();

#define RVB_REQUIRE(cond, ...)                              \
  do {                                                      \
    if (!(cond)) {                                          \
      ::rvb::set_error(__VA_ARGS__);                        \
      return RVB_ERR_ARG;                                   \
    }                                                       \
  } while (0)

#define RVB_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::rvb::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                  \
      return RVB_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

inline bool epilogue_ok(int e) {
  const int epi = e & 0xf;
  return (e & ~(0xf | RVB_EPI_TIME_MAJOR)) == 0 && epi >= RVB_EPI_POWER && epi <= RVB_EPI_POWER_P &&
         !((e & RVB_EPI_TIME_MAJOR) && epi == RVB_EPI_COMPLEX);
}

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---- warp reductions (butterfly: every lane ends with the same bits) ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ unsigned warp_max_u32(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// a / b for a divisor shared by a whole row: q = a * (1/b), one residual correction.  With the correctly rounded
// reciprocal this is the correctly rounded quotient except for a vanishing set of operands (Markstein), where it is
// one ulp off -- three instructions instead of the ~10 of the IEEE sequence with its slow-path check, which made
// these kernels issue-bound (16 M warp instructions for 20 480 VAT rows, profiles/r01e).  Rows whose divisor is 0, inf,
// NaN or so small that 1/b overflows take the real division (kFast = false) under a warp-uniform branch.
__device__ __forceinline__ bool rcp_usable(float rcp_b) { return fabsf(rcp_b) <= 3.0e38f && rcp_b != 0.f; }
template <bool kFast>
__device__ __forceinline__ float div_by(float a, float b, float rcp_b) {
  if constexpr (kFast) {
    const float q = a * rcp_b;
    return fmaf(fmaf(-q, b, a), rcp_b, q);
  } else {
    return a / b;
  }
}

// torch.clamp(v, 0, 1): NaN propagates (fminf/fmaxf would drop it).
__device__ __forceinline__ float clamp01(float v) { return v < 0.f ? 0.f : (v > 1.f ? 1.f : v); }

// std::max(v, lo) / torch.clamp(v, lo, hi) as ATen evaluates them: a NaN in `v` comes out as NaN.
__device__ __forceinline__ float max_nan(float v, float lo) { return v < lo ? lo : v; }
__device__ __forceinline__ float clamp_nan(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Monotone float -> uint32 key (works with a zero-initialised atomicMax target:
// every finite/inf float maps above 0).  NaN (positive payload) maps above +inf.
__device__ __forceinline__ unsigned f2key(float f) {
  unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
  unsigned b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(b);
}

// Round-to-nearest fp32 -> tf32 (10-bit mantissa, low 13 bits cleared).
__device__ __forceinline__ float to_tf32(float v) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// ---- STFT output formats, shared by the tcgen05 contractions and the single-bin kernels ----
// Single-float formats (power / magnitude / phase / power_p) may carry RVB_EPI_TIME_MAJOR:
//   bin-major  out0[(b*n_out_bins + k)*T + t]      what STFT.forward returns
//   time-major out0[(b*T + t)*n_out_bins + k]      what the Mel kernel reads (16-byte stores per thread)
// RVB_EPI_POWER is evaluated as re^2 + im^2: the reference's sqrt followed by **2.0 (Spectrogram.py:227,231,458)
// returns that value to within 1 ulp, and the IEEE sqrt sequence (MUFU + Newton + slow-path branch, x128 per
// thread and tile) made the tensor-core epilogue instruction-fetch bound (profiles/r01c).
static __device__ __noinline__ float stft_value_slow(int epi, float power, float re, float im) {
  if (epi == RVB_EPI_PHASE) return atan2f(-im + 0.0f, re);             // Spectrogram.py:237
  return powf(sqrtf(re * re + im * im), power);                         // general `power` (Spectrogram.py:458)
}
template <int kEpi>
__device__ __forceinline__ float stft_value_t(float power, float re, float im) {
  if constexpr (kEpi == RVB_EPI_POWER) return fmaf(re, re, im * im);
  else if constexpr (kEpi == RVB_EPI_MAGNITUDE) return sqrtf(re * re + im * im);   // Spectrogram.py:227,231
  else return stft_value_slow(kEpi, power, re, im);
}
__device__ __forceinline__ float stft_value(int epi, float power, float re, float im) {
  if (epi == RVB_EPI_POWER) return stft_value_t<RVB_EPI_POWER>(power, re, im);
  if (epi == RVB_EPI_MAGNITUDE) return stft_value_t<RVB_EPI_MAGNITUDE>(power, re, im);
  return stft_value_slow(epi, power, re, im);
}
__device__ __forceinline__ void stft_store(int epilogue, float power, float re, float im, float* __restrict__ out0,
                                           int b, int k, int t, int n_out_bins, int n_frames) {
  const int epi = epilogue & 0xf;
  if (epi == RVB_EPI_COMPLEX) {
    reinterpret_cast<float2*>(out0)[((int64_t)b * n_out_bins + k) * n_frames + t] = make_float2(re, -im);   // :234
  } else if (epilogue & RVB_EPI_TIME_MAJOR) {
    out0[((int64_t)b * n_frames + t) * n_out_bins + k] = stft_value(epi, power, re, im);
  } else {
    out0[((int64_t)b * n_out_bins + k) * n_frames + t] = stft_value(epi, power, re, im);
  }
}

// Epilogue of one accumulator chunk: 32 consecutive bins [k0, k0+32) of frame (b, t) held by one thread.
// `scale` undoes the power-of-two operand scaling of the fp16 contraction (1 for tf32 operands: exact either way).
// One compact code path per (format, layout): the dispatch is per chunk, never per value.
template <int kEpi, bool kTimeMajor>
__device__ __forceinline__ void stft_store_chunk_t(float power, const uint32_t (&re)[32], const uint32_t (&im)[32],
                                                   float scale, float re_add, float* __restrict__ out0, int b, int k0,
                                                   int t, int n_out_bins, int n_store_bins, int n_frames) {
  if constexpr (kTimeMajor) {
    float* row = out0 + ((int64_t)b * n_frames + t) * n_out_bins + k0;
    if (k0 + 32 <= n_store_bins && (n_out_bins & 3) == 0) {
      float4* dst = reinterpret_cast<float4*>(row);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 v;
        v.x = stft_value_t<kEpi>(power, fmaf(__uint_as_float(re[4 * i + 0]), scale, re_add), __uint_as_float(im[4 * i + 0]) * scale);
        v.y = stft_value_t<kEpi>(power, fmaf(__uint_as_float(re[4 * i + 1]), scale, re_add), __uint_as_float(im[4 * i + 1]) * scale);
        v.z = stft_value_t<kEpi>(power, fmaf(__uint_as_float(re[4 * i + 2]), scale, re_add), __uint_as_float(im[4 * i + 2]) * scale);
        v.w = stft_value_t<kEpi>(power, fmaf(__uint_as_float(re[4 * i + 3]), scale, re_add), __uint_as_float(im[4 * i + 3]) * scale);
        dst[i] = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (k0 + i < n_store_bins)
          row[i] = stft_value_t<kEpi>(power, fmaf(__uint_as_float(re[i]), scale, re_add), __uint_as_float(im[i]) * scale);
    }
  } else {
    const int64_t base = ((int64_t)b * n_out_bins + k0) * n_frames + t;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (k0 + i < n_store_bins) {
        const float r = fmaf(__uint_as_float(re[i]), scale, re_add), m = __uint_as_float(im[i]) * scale;
        if constexpr (kEpi == RVB_EPI_COMPLEX)
          reinterpret_cast<float2*>(out0)[base + (int64_t)i * n_frames] = make_float2(r, -m);    // Spectrogram.py:234
        else
          out0[base + (int64_t)i * n_frames] = stft_value_t<kEpi>(power, r, m);
      }
    }
  }
}

__device__ __forceinline__ void stft_store_chunk(int epilogue, float power, const uint32_t (&re)[32],
                                                 const uint32_t (&im)[32], float scale, float re_add,
                                                 float* __restrict__ out0, int b, int k0, int t, int n_out_bins,
                                                 int n_store_bins, int n_frames) {
#define RVB_EPI_CASE(E, TM)                                                                                         \
  case (E) | ((TM) ? RVB_EPI_TIME_MAJOR : 0):                                                                       \
    stft_store_chunk_t<E, TM>(power, re, im, scale, re_add, out0, b, k0, t, n_out_bins, n_store_bins, n_frames);   \
    break;
  switch (epilogue) {
    RVB_EPI_CASE(RVB_EPI_POWER, true)
    RVB_EPI_CASE(RVB_EPI_POWER, false)
    RVB_EPI_CASE(RVB_EPI_MAGNITUDE, true)
    RVB_EPI_CASE(RVB_EPI_MAGNITUDE, false)
    RVB_EPI_CASE(RVB_EPI_COMPLEX, false)
    RVB_EPI_CASE(RVB_EPI_PHASE, true)
    RVB_EPI_CASE(RVB_EPI_PHASE, false)
    RVB_EPI_CASE(RVB_EPI_POWER_P, true)
    RVB_EPI_CASE(RVB_EPI_POWER_P, false)
    default: break;
  }
#undef RVB_EPI_CASE
}

}  // namespace rvb
