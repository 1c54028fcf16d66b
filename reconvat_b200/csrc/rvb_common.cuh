// Shared device/host helpers for the reconvat_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "../../include/rvb.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "reconvat_b200 kernels are written for sm_100a (B200) only"
#endif

namespace rvb {

// ---- host-side error plumbing (thread-local last-error string, see rvb_last_error) ----
void set_error(const char* fmt, ...);
int check_launch(const char* what);           // cudaGetLastError -> RVB_ERR_LAUNCH

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

// torch.clamp(v, 0, 1): NaN propagates (fminf/fmaxf would drop it).
__device__ __forceinline__ float clamp01(float v) { return v < 0.f ? 0.f : (v > 1.f ? 1.f : v); }

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

// ---- STFT output formats, shared by the tcgen05 contraction and the single-bin kernel ----
__device__ __forceinline__ void stft_store(int epilogue, float power, float re, float im, float* __restrict__ out0,
                                           int64_t idx /* (b*n_out_bins + k)*T + t */) {
  if (epilogue == RVB_EPI_COMPLEX) {
    reinterpret_cast<float2*>(out0)[idx] = make_float2(re, -im);        // Spectrogram.py:234
  } else if (epilogue == RVB_EPI_PHASE) {
    out0[idx] = atan2f(-im + 0.0f, re);                                 // Spectrogram.py:237
  } else {
    float mag = sqrtf(re * re + im * im);                               // Spectrogram.py:227,231
    if (epilogue == RVB_EPI_POWER) mag = mag * mag;                     // **2.0 (Spectrogram.py:458)
    else if (epilogue == RVB_EPI_POWER_P) mag = powf(mag, power);
    out0[idx] = mag;
  }
}


}  // namespace rvb
