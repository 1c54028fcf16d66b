// VAT perturbation loop kernels (HBM-bound, one warp per 229-float Mel row).
//
// Reference arithmetic: model/self_attention_VAT.py:162-202 (UNet_VAT.forward) and its siblings
// model/UNet_onset.py:116-162, model/onset_frame_VAT.py:175-207, model/VAT.py:20-40.
// Layout: x, d, g, r_adv, x_adv, d_hat are [n_rows][row_len] contiguous fp32, n_rows = B*T.
// A row (916 B for 229 mels) is only 4-byte aligned, so each lane owns elements lane, lane+32, ...
// (every warp-wide access is one fully coalesced 128-byte request) and keeps them in registers:
// each tensor is read or written exactly once.
#include "rvb_common.cuh"

namespace rvb {

extern void count_launch();

constexpr int kRowsPerBlock = 8;     // 8 warps / 256 threads
constexpr int kMaxPerLane = 16;      // rows up to 512 floats

template <int NPL>
struct RowRegs {
  float v[NPL];
};

template <int NPL>
__device__ __forceinline__ void load_row(RowRegs<NPL>& r, const float* __restrict__ p, int row_len, int lane) {
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    int c = lane + i * kWarp;
    r.v[i] = (c < row_len) ? __ldg(p + c) : 0.f;
  }
}
template <int NPL>
__device__ __forceinline__ void store_row(const RowRegs<NPL>& r, float* __restrict__ p, int row_len, int lane) {
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    int c = lane + i * kWarp;
    if (c < row_len) p[c] = r.v[i];
  }
}
template <int NPL>
__device__ __forceinline__ float row_sumsq(const RowRegs<NPL>& r) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) s = fmaf(r.v[i], r.v[i], s);
  return warp_sum(s);
}

// ---- V1: x_adv = clamp(x + xi * d/||d||) -------------------------------------------------
template <int NPL>
__global__ void __launch_bounds__(kRowsPerBlock* kWarp)
vat_perturb_kernel(const float* __restrict__ x, const float* __restrict__ d, float* __restrict__ x_adv,
                   int64_t n_rows, int row_len, float xi, int do_clamp) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int64_t off = row * row_len;
  RowRegs<NPL> rd, rx;
  load_row(rd, d + off, row_len, lane);
  load_row(rx, x + off, row_len, lane);
  const float n = sqrtf(row_sumsq(rd));       // torch.norm(d, dim=-1)
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    float s = rx.v[i] + xi * (rd.v[i] / n);   // x + XI * (d / n)
    rx.v[i] = do_clamp ? clamp01(s) : s;
  }
  store_row(rx, x_adv + off, row_len, lane);
}

// ---- V3: power-iteration backward + finalisation ----------------------------------------
template <int NPL>
__device__ __forceinline__ void finalize_row(RowRegs<NPL>& dp /* in: d' ; out: dhat' */, RowRegs<NPL>& rx,
                                             float* __restrict__ r_adv, float* __restrict__ x_adv,
                                             float* __restrict__ d_hat, int64_t off, int row_len, int lane,
                                             float eps, int do_clamp, int32_t* status_flag) {
  const float n2 = sqrtf(row_sumsq(dp));
  RowRegs<NPL> rr;
  unsigned bad = 0;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const bool in = lane + i * kWarp < row_len;
    float dh = dp.v[i] / n2;                  // _l2_normalize(d)
    float r = eps * dh;                       // r_adv
    if (in) bad |= (isnan(r) ? 1u : 0u) | (isinf(r) ? 2u : 0u);
    float s = rx.v[i] + r;
    dp.v[i] = dh;
    rr.v[i] = r;
    rx.v[i] = do_clamp ? clamp01(s) : s;
  }
  store_row(rr, r_adv + off, row_len, lane);
  store_row(rx, x_adv + off, row_len, lane);
  store_row(dp, d_hat + off, row_len, lane);
  bad = __reduce_or_sync(kFull, bad);
  if (bad && lane == 0 && status_flag) atomicOr(status_flag, (int)bad);
}

template <int NPL>
__global__ void __launch_bounds__(kRowsPerBlock* kWarp)
vat_finalize_kernel(const float* __restrict__ g, const float* __restrict__ d, const float* __restrict__ x,
                    float* __restrict__ r_adv, float* __restrict__ x_adv, float* __restrict__ d_hat,
                    int64_t n_rows, int row_len, float xi, float eps, float scale, int do_clamp,
                    int32_t* status_flag) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int64_t off = row * row_len;
  RowRegs<NPL> rd, rx, rg;
  load_row(rd, d + off, row_len, lane);
  load_row(rx, x + off, row_len, lane);
  load_row(rg, g + off, row_len, lane);
  const float n = sqrtf(row_sumsq(rd));
  // gd = xi * g * [0 <= x + xi*d/n <= 1]   (clamp passes the gradient on the closed interval)
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    float gm = rg.v[i];
    if (do_clamp) {
      float s = rx.v[i] + xi * (rd.v[i] / n);
      gm = (s >= 0.f && s <= 1.f) ? gm : 0.f;
    }
    float gd = xi * gm;
    rg.v[i] = gd;
    dot = fmaf(gd, rd.v[i], dot);
  }
  dot = warp_sum(dot);
  const float c = dot / (n * n * n);
#pragma unroll
  for (int i = 0; i < NPL; ++i) rd.v[i] = (rg.v[i] / n - rd.v[i] * c) * scale;   // d.grad * scale
  finalize_row(rd, rx, r_adv, x_adv, d_hat, off, row_len, lane, eps, do_clamp, status_flag);
}

template <int NPL>
__global__ void __launch_bounds__(kRowsPerBlock* kWarp)
vat_direct_kernel(const float* __restrict__ d, const float* __restrict__ x, float* __restrict__ r_adv,
                  float* __restrict__ x_adv, float* __restrict__ d_hat, int64_t n_rows, int row_len, float eps,
                  int do_clamp, int32_t* status_flag) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int64_t off = row * row_len;
  RowRegs<NPL> rd, rx;
  load_row(rd, d + off, row_len, lane);
  load_row(rx, x + off, row_len, lane);
  finalize_row(rd, rx, r_adv, x_adv, d_hat, off, row_len, lane, eps, do_clamp, status_flag);
}

// ---- V2: d mean-BCE / d p ---------------------------------------------------------------
__device__ __forceinline__ float bce_grad_one(float p, float y, float s) {
  return (p - y) / fmaxf((1.f - p) * p, 1e-12f) * s;
}

__global__ void __launch_bounds__(256)
bce_grad_kernel(const float* __restrict__ p, const float* __restrict__ y, float* __restrict__ grad, int64_t n,
                const float* __restrict__ gscale_dev, float gscale, int vec_ok) {
  const float s = (gscale_dev ? __ldg(gscale_dev) : 1.f) * gscale / (float)n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    const float4* p4 = reinterpret_cast<const float4*>(p);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    float4* g4 = reinterpret_cast<float4*>(grad);
    for (int64_t j = i; j < n4; j += stride) {
      float4 a = __ldg(p4 + j), b = __ldg(y4 + j), o;
      o.x = bce_grad_one(a.x, b.x, s);
      o.y = bce_grad_one(a.y, b.y, s);
      o.z = bce_grad_one(a.z, b.z, s);
      o.w = bce_grad_one(a.w, b.w, s);
      g4[j] = o;
    }
    for (int64_t j = (n4 << 2) + i; j < n; j += stride) grad[j] = bce_grad_one(__ldg(p + j), __ldg(y + j), s);
  } else {
    for (int64_t j = i; j < n; j += stride) grad[j] = bce_grad_one(__ldg(p + j), __ldg(y + j), s);
  }
}

// ---- V4: mean BCE, deterministic ---------------------------------------------------------
__device__ __forceinline__ float bce_one(float p, float y) {
  // ATen: (y - 1) * max(log1p(-p), -100) - y * max(log(p), -100)
  return (y - 1.f) * fmaxf(log1pf(-p), -100.f) - y * fmaxf(logf(p), -100.f);
}

constexpr int kBceMaxBlocks = 1024;

__global__ void __launch_bounds__(256)
bce_mean_kernel(const float* __restrict__ p, const float* __restrict__ y, int64_t n, float* __restrict__ loss,
                float* __restrict__ workspace, int vec_ok) {
  __shared__ float warp_part[8];
  __shared__ bool is_last;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    const float4* p4 = reinterpret_cast<const float4*>(p);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    for (int64_t j = i; j < n4; j += stride) {
      float4 a = __ldg(p4 + j), b = __ldg(y4 + j);
      acc += (bce_one(a.x, b.x) + bce_one(a.y, b.y)) + (bce_one(a.z, b.z) + bce_one(a.w, b.w));
    }
    for (int64_t j = (n4 << 2) + i; j < n; j += stride) acc += bce_one(__ldg(p + j), __ldg(y + j));
  } else {
    for (int64_t j = i; j < n; j += stride) acc += bce_one(__ldg(p + j), __ldg(y + j));
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  unsigned* ticket = reinterpret_cast<unsigned*>(workspace + kBceMaxBlocks);
  if (threadIdx.x == 0) {
    float b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) b += warp_part[w];
    workspace[blockIdx.x] = b;
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    // fixed-order final sum in double: the result does not depend on block scheduling
    double s = 0.0;
    for (int j = threadIdx.x; j < (int)gridDim.x; j += blockDim.x) s += (double)__ldcg(workspace + j);
    __shared__ double dsum[256];
    dsum[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) dsum[threadIdx.x] += dsum[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      *loss = (float)(dsum[0] / (double)n);
      *ticket = 0u;   // self-cleaning for the next call on this stream
    }
  }
}

template <typename F>
static int dispatch_npl(int row_len, F&& f) {
  if (row_len <= 32 * 4) return f(std::integral_constant<int, 4>{});
  if (row_len <= 32 * 8) return f(std::integral_constant<int, 8>{});
  if (row_len <= 32 * kMaxPerLane) return f(std::integral_constant<int, kMaxPerLane>{});
  set_error("row_len %d > %d not supported", row_len, 32 * kMaxPerLane);
  return RVB_ERR_ARG;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace rvb

using namespace rvb;

extern "C" int rvb_vat_perturb(const float* x, const float* d, float* x_adv, int64_t n_rows, int row_len, float xi,
                               int do_clamp, rvb_stream_t stream) {
  RVB_REQUIRE(x && d && x_adv, "rvb_vat_perturb: null pointer");
  RVB_REQUIRE(n_rows >= 0 && row_len > 0, "rvb_vat_perturb: bad shape (%lld, %d)", (long long)n_rows, row_len);
  if (n_rows == 0) return RVB_OK;
  const unsigned grid = (unsigned)((n_rows + kRowsPerBlock - 1) / kRowsPerBlock);
  return dispatch_npl(row_len, [&](auto npl) {
    vat_perturb_kernel<decltype(npl)::value><<<grid, kRowsPerBlock * kWarp, 0, (cudaStream_t)stream>>>(
        x, d, x_adv, n_rows, row_len, xi, do_clamp);
    count_launch();
    return check_launch("vat_perturb_kernel");
  });
}

extern "C" int rvb_vat_finalize(const float* g, const float* d, const float* x, float* r_adv, float* x_adv,
                                float* d_hat, int64_t n_rows, int row_len, float xi, float eps, float scale,
                                int do_clamp, int32_t* status_flag, rvb_stream_t stream) {
  RVB_REQUIRE(g && d && x && r_adv && x_adv && d_hat, "rvb_vat_finalize: null pointer");
  RVB_REQUIRE(n_rows >= 0 && row_len > 0, "rvb_vat_finalize: bad shape (%lld, %d)", (long long)n_rows, row_len);
  if (n_rows == 0) return RVB_OK;
  const unsigned grid = (unsigned)((n_rows + kRowsPerBlock - 1) / kRowsPerBlock);
  return dispatch_npl(row_len, [&](auto npl) {
    vat_finalize_kernel<decltype(npl)::value><<<grid, kRowsPerBlock * kWarp, 0, (cudaStream_t)stream>>>(
        g, d, x, r_adv, x_adv, d_hat, n_rows, row_len, xi, eps, scale, do_clamp, status_flag);
    count_launch();
    return check_launch("vat_finalize_kernel");
  });
}

extern "C" int rvb_vat_direct(const float* d, const float* x, float* r_adv, float* x_adv, float* d_hat,
                              int64_t n_rows, int row_len, float eps, int do_clamp, int32_t* status_flag,
                              rvb_stream_t stream) {
  RVB_REQUIRE(d && x && r_adv && x_adv && d_hat, "rvb_vat_direct: null pointer");
  RVB_REQUIRE(n_rows >= 0 && row_len > 0, "rvb_vat_direct: bad shape (%lld, %d)", (long long)n_rows, row_len);
  if (n_rows == 0) return RVB_OK;
  const unsigned grid = (unsigned)((n_rows + kRowsPerBlock - 1) / kRowsPerBlock);
  return dispatch_npl(row_len, [&](auto npl) {
    vat_direct_kernel<decltype(npl)::value><<<grid, kRowsPerBlock * kWarp, 0, (cudaStream_t)stream>>>(
        d, x, r_adv, x_adv, d_hat, n_rows, row_len, eps, do_clamp, status_flag);
    count_launch();
    return check_launch("vat_direct_kernel");
  });
}

static unsigned flat_grid(int64_t n, int per_thread) {
  int64_t blocks = (n + 256LL * per_thread - 1) / (256LL * per_thread);
  // a few waves of 148 SMs x 8 resident blocks is plenty for a streaming kernel
  const int64_t cap = 148 * 16;
  return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

extern "C" int rvb_bce_grad(const float* p, const float* y, float* grad, int64_t n, const float* gscale_dev,
                            float gscale, rvb_stream_t stream) {
  RVB_REQUIRE(p && y && grad, "rvb_bce_grad: null pointer");
  RVB_REQUIRE(n >= 0, "rvb_bce_grad: negative size");
  if (n == 0) return RVB_OK;
  const int vec = aligned16(p) && aligned16(y) && aligned16(grad);
  bce_grad_kernel<<<flat_grid(n, 8), 256, 0, (cudaStream_t)stream>>>(p, y, grad, n, gscale_dev, gscale, vec);
  count_launch();
  return check_launch("bce_grad_kernel");
}

extern "C" int rvb_bce_mean(const float* p, const float* y, int64_t n, float* loss, float* workspace,
                            rvb_stream_t stream) {
  RVB_REQUIRE(p && y && loss && workspace, "rvb_bce_mean: null pointer");
  RVB_REQUIRE(n > 0, "rvb_bce_mean: empty input (the reference returns NaN for an empty mean)");
  unsigned grid = flat_grid(n, 16);
  if (grid > (unsigned)kBceMaxBlocks) grid = kBceMaxBlocks;
  const int vec = aligned16(p) && aligned16(y);
  bce_mean_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, y, n, loss, workspace, vec);
  count_launch();
  return check_launch("bce_mean_kernel");
}
