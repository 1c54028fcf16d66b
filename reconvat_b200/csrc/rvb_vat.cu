// VAT perturbation loop kernels (HBM-bound, one warp per 229-float Mel row).
//
// Reference arithmetic: model/self_attention_VAT.py:162-202 (UNet_VAT.forward) and its siblings
// model/UNet_onset.py:116-162, model/onset_frame_VAT.py:175-207, model/VAT.py:20-40.
// Layout: x, d, g, r_adv, x_adv, d_hat are [n_rows][row_len] contiguous fp32, n_rows = B*T.
// A row (916 B for 229 mels) is only 4-byte aligned, so each lane owns elements lane, lane+32, ...
// (every warp-wide access is one fully coalesced 128-byte request) and keeps them in registers:
// each tensor is read or written exactly once.
#include <type_traits>

#include <curand_kernel.h>        // device-side Philox4x32-10 and Box-Muller, the functions ATen's normal_ kernel calls

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
// Persistent: the grid is a few blocks per SM and every warp walks rows with a fixed stride, the loads of its NEXT row
// issued before the current row is reduced -- a warp always has 2 x NPL x 128 bytes in flight.  One-shot blocks (one row
// per warp, then exit) ramp up and drain once per 8 rows; beside a resident contraction CTA of another stream, where
// only one such block fits per SM, that left the memory pipe idle between blocks (144 -> 217 us for the HBM kernels of a
// step, profiles/r01f_experiments.md).
constexpr int kPersistBlocksPerSM = 3;

template <int NPL>
__global__ void __launch_bounds__(kRowsPerBlock* kWarp) __maxnreg__(NPL <= 8 ? 72 : 128)
vat_perturb_kernel(const float* __restrict__ x, const float* __restrict__ d, float* __restrict__ x_adv,
                   int64_t n_rows, int row_len, float xi, int do_clamp) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * kRowsPerBlock;
  int64_t row = (int64_t)blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  RowRegs<NPL> nd, nx;
  load_row(nd, d + row * row_len, row_len, lane);
  load_row(nx, x + row * row_len, row_len, lane);
  for (;;) {
    RowRegs<NPL> rd = nd, rx = nx;
    const int64_t off = row * row_len, next = row + stride;
    if (next < n_rows) {
      load_row(nd, d + next * row_len, row_len, lane);
      load_row(nx, x + next * row_len, row_len, lane);
    }
    const float n = sqrtf(row_sumsq(rd));       // torch.norm(d, dim=-1)
    const float rn = 1.f / n;
    auto body = [&](auto fast) {
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        float s = rx.v[i] + xi * div_by<decltype(fast)::value>(rd.v[i], n, rn);   // x + XI * (d / n)
        rx.v[i] = do_clamp ? clamp01(s) : s;
      }
    };
    if (rcp_usable(rn)) body(std::true_type{}); else body(std::false_type{});
    store_row(rx, x_adv + off, row_len, lane);
    if (next >= n_rows) break;
    row = next;
  }
}

// ---- V0 + V1: draw d ~ N(0, 1) IN the perturb kernel, bit-identical to torch.randn_like ------------------------
// ATen's normal_ on a contiguous float tensor of n elements (aten/src/ATen/native/cuda/DistributionTemplates.h):
// grid = min(#SM * (maxThreadsPerSM / 256), ceil(n / 256)) blocks of 256 threads, TT = 256 * grid threads in total;
// thread idx owns the Philox4x32-10 stream (key = seed, subsequence = idx, offset = the generator's offset) and its
// j-th curand_normal4 call fills elements idx + TT * (4 j + ii), ii = 0 .. 3; afterwards the generator's offset has
// advanced by ((n - 1) / (4 TT) + 1) * 4.  So element li is component ii of Box-Muller on Philox(key,
// counter = (offset / 4 + j, idx)) with q = li / TT, idx = li % TT, j = q >> 2, ii = q & 3 -- computable for any
// element on its own.  A row-wise kernel cannot share a Philox call between its four elements (they lie TT apart), so
// each element pays a full call: ~95 instructions, the price of bit parity (the separate ATen kernel it replaces cost
// 13 us per step, one launch and a 19 MB write + read of d).
struct DrawArgs {
  unsigned long long seed, offset;      // used when dev_state == nullptr (eager: read from torch's generator)
  unsigned long long* dev_state;        // [seed, offset, ticket]: CUDA-graph replays carry their own stream
  unsigned long long increment;         // what ATen would add to the offset
  unsigned tt;                          // TT
};

__device__ __forceinline__ float draw_normal(unsigned long long seed, unsigned long long ctr_lo, unsigned idx, unsigned ii) {
  const uint4 ctr = make_uint4((unsigned)ctr_lo, (unsigned)(ctr_lo >> 32), idx, 0u);
  const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
  const uint4 r = curand_Philox4x32_10(ctr, key);
  const float2 n2 = _curand_box_muller((ii & 2) ? r.z : r.x, (ii & 2) ? r.w : r.y);
  return (ii & 1) ? n2.y : n2.x;
}

// The same draw as a kernel of its own, with ATen's amortisation: thread <-> Philox stream idx, one Philox call and two
// Box-Muller evaluations per FOUR elements (idx + TT (4 j + ii)), every store warp-coalesced.  ATen's own kernel spends
// 13 us on the 4.7 M elements of a B = 32 step (four dependent loop trips with a __syncthreads each); this one issues
// all trips of a thread at once.  What the fused row kernel above cannot do -- share a call between its elements --
// costs it 4x the Philox work, which matters when the kernel runs beside the tensor-bound contraction of another
// stream: there issue slots are scarce and HBM bytes are cheap.
template <int kTrips>
__global__ void __launch_bounds__(256)
randn_like_kernel(float* __restrict__ out, int64_t n, DrawArgs a) {
  unsigned long long seed = a.seed, offset = a.offset;
  if (a.dev_state) { seed = a.dev_state[0]; offset = a.dev_state[1]; }
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < a.tt) {
    const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
    const int64_t n_trips = (n + 4ll * a.tt - 1) / (4ll * a.tt);
    for (int64_t j0 = 0; j0 < n_trips; j0 += kTrips) {
      float4 v[kTrips];
#pragma unroll
      for (int t = 0; t < kTrips; ++t) {
        const unsigned long long c = (offset >> 2) + (unsigned long long)(j0 + t);
        const uint4 r = curand_Philox4x32_10(make_uint4((unsigned)c, (unsigned)(c >> 32), idx, 0u), key);
        const float2 lo = _curand_box_muller(r.x, r.y), hi = _curand_box_muller(r.z, r.w);
        v[t] = make_float4(lo.x, lo.y, hi.x, hi.y);
      }
#pragma unroll
      for (int t = 0; t < kTrips; ++t) {
        const int64_t base = (int64_t)idx + 4ll * a.tt * (j0 + t);
        if (base < n) out[base] = v[t].x;
        if (base + a.tt < n) out[base + a.tt] = v[t].y;
        if (base + 2ll * a.tt < n) out[base + 2ll * a.tt] = v[t].z;
        if (base + 3ll * a.tt < n) out[base + 3ll * a.tt] = v[t].w;
      }
    }
  }
  if (a.dev_state) {
    __shared__ bool is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      is_last = atomicAdd(a.dev_state + 2, 1ull) == (unsigned long long)gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
      a.dev_state[1] = offset + a.increment;
      a.dev_state[2] = 0ull;
      __threadfence();
    }
  }
}

template <int NPL>
__global__ void __launch_bounds__(kRowsPerBlock* kWarp)
vat_perturb_draw_kernel(const float* __restrict__ x, float* __restrict__ d_out, float* __restrict__ x_adv,
                        int64_t n_rows, int row_len, float xi, int do_clamp, DrawArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  unsigned long long seed = a.seed, offset = a.offset;
  if (a.dev_state) { seed = a.dev_state[0]; offset = a.dev_state[1]; }
  if (row < n_rows) {
    const int64_t off = row * row_len;
    // (q, idx) of the row's first element; a row crosses at most one multiple of TT
    const unsigned long long q0 = (unsigned long long)off / a.tt;
    const unsigned rem0 = (unsigned)((unsigned long long)off - q0 * a.tt);
    RowRegs<NPL> rd, rx;
    load_row(rx, x + off, row_len, lane);
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + i * kWarp;
      float v = 0.f;
      if (c < row_len) {
        unsigned idx = rem0 + (unsigned)c;
        unsigned long long q = q0;
        if (idx >= a.tt) { idx -= a.tt; ++q; }
        v = draw_normal(seed, (offset >> 2) + (q >> 2), idx, (unsigned)(q & 3));
      }
      rd.v[i] = v;
    }
    if (d_out) store_row(rd, d_out + off, row_len, lane);
    const float n = sqrtf(row_sumsq(rd));       // torch.norm(d, dim=-1)
    const float rn = 1.f / n;
    auto body = [&](auto fast) {
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        float s = rx.v[i] + xi * div_by<decltype(fast)::value>(rd.v[i], n, rn);   // x + XI * (d / n)
        rx.v[i] = do_clamp ? clamp01(s) : s;
      }
    };
    if (rcp_usable(rn)) body(std::true_type{}); else body(std::false_type{});
    store_row(rx, x_adv + off, row_len, lane);
  }
  if (a.dev_state) {
    // the last block to finish advances the device-resident offset (every block has read it by then)
    __shared__ bool is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      is_last = atomicAdd(a.dev_state + 2, 1ull) == (unsigned long long)gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
      a.dev_state[1] = offset + a.increment;
      a.dev_state[2] = 0ull;
      __threadfence();
    }
  }
}

// ---- V3: power-iteration backward + finalisation ----------------------------------------
// Per-row result of the finalisation: NaN / Inf bits of r_adv and sum |dhat| (the `r_norm.abs().mean()` the reference's
// run_on_batch logs, model/self_attention_VAT.py:1149).
struct RowStat {
  unsigned bad;
  float abs_sum;
};

template <int NPL>
__device__ __forceinline__ RowStat finalize_row(RowRegs<NPL>& dp /* in: d' ; out: dhat' */, RowRegs<NPL>& rx,
                                                float* __restrict__ r_adv, float* __restrict__ x_adv,
                                                float* __restrict__ d_hat, int64_t off, int row_len, int lane,
                                                float eps, int do_clamp) {
  const float n2 = sqrtf(row_sumsq(dp));
  const float rn2 = 1.f / n2;
  RowRegs<NPL> rr;
  unsigned bad = 0;
  float asum = 0.f;
  // NaN / Inf in r_adv = eps d'/||d'|| needs a NaN, an Inf or an all-zero row in d' -- and each of those makes the
  // reciprocal of the norm unusable (0, Inf or NaN).  Rows on the fast path (usable reciprocal: every d' finite, norm
  // finite and positive) cannot raise a bit, so only the slow path looks at the elements.  (Padding lanes hold 0.)
  auto body = [&](auto fast) {
    constexpr bool kFast = decltype(fast)::value;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      float dh = div_by<kFast>(dp.v[i], n2, rn2);                      // _l2_normalize(d)
      float r = eps * dh;                     // r_adv
      if constexpr (kFast) {
        asum += fabsf(dh);
      } else if (lane + i * kWarp < row_len) {
        bad |= (isnan(r) ? 1u : 0u) | (isinf(r) ? 2u : 0u);
        asum += fabsf(dh);
      }
      float s = rx.v[i] + r;
      dp.v[i] = dh;
      rr.v[i] = r;
      rx.v[i] = do_clamp ? clamp01(s) : s;
    }
  };
  if (rcp_usable(rn2)) body(std::true_type{}); else body(std::false_type{});
  store_row(rr, r_adv + off, row_len, lane);
  store_row(rx, x_adv + off, row_len, lane);
  if (d_hat) store_row(dp, d_hat + off, row_len, lane);    // stats flavour: optional (its mean |.| is reduced here)
  RowStat st;
  st.bad = __reduce_or_sync(kFull, bad);
  st.abs_sum = warp_sum(asum);
  return st;
}

// Where the per-row results go.  Plain flavour: OR the bits into a flag the caller zeroed.  Stats flavour (workspace
// != nullptr): per-block partials, and the LAST block to finish adds them in a fixed order (double), WRITES the flag
// and the mean |dhat| -- nobody has to zero the flag, and the logged metric costs no extra pass over d_hat.
// workspace: [n_blocks] float partial sums | [n_blocks] uint32 bits | uint32 ticket (zero once; self-cleaning).
__device__ __forceinline__ void finalize_block_stats(RowStat st, bool valid, int32_t* __restrict__ status_flag,
                                                     float* __restrict__ dhat_abs_mean, float* __restrict__ workspace,
                                                     double n_elems) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (!workspace) {
    if (valid && st.bad && lane == 0 && status_flag) atomicOr(status_flag, (int)st.bad);
    return;
  }
  __shared__ float s_abs[kRowsPerBlock];
  __shared__ unsigned s_bad[kRowsPerBlock];
  __shared__ bool is_last;
  if (lane == 0) { s_abs[warp] = valid ? st.abs_sum : 0.f; s_bad[warp] = valid ? st.bad : 0u; }
  __syncthreads();
  float* part = workspace;
  unsigned* bits = reinterpret_cast<unsigned*>(workspace) + gridDim.x;
  unsigned* ticket = bits + gridDim.x;
  if (threadIdx.x == 0) {
    float a = 0.f;
    unsigned bd = 0;
#pragma unroll
    for (int w = 0; w < kRowsPerBlock; ++w) { a += s_abs[w]; bd |= s_bad[w]; }
    part[blockIdx.x] = a;
    bits[blockIdx.x] = bd;
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double s = 0.0;
  unsigned bd = 0;
  for (int j = threadIdx.x; j < (int)gridDim.x; j += blockDim.x) { s += (double)__ldcg(part + j); bd |= __ldcg(bits + j); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
  bd = __reduce_or_sync(kFull, bd);
  __shared__ double d_part[kRowsPerBlock];
  __syncthreads();                                         // s_bad is reused below
  if (lane == 0) { d_part[warp] = s; s_bad[warp] = bd; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    unsigned all = 0;
#pragma unroll
    for (int w = 0; w < kRowsPerBlock; ++w) { tot += d_part[w]; all |= s_bad[w]; }
    if (dhat_abs_mean) *dhat_abs_mean = (float)(tot / n_elems);
    if (status_flag) *status_flag = (int)all;
    *ticket = 0u;
  }
}

// Persistent like vat_perturb_kernel: a warp walks rows with a fixed stride, the three loads of its next row in flight
// while the current row goes through its three reductions; the per-row results (NaN / Inf bits, sum |dhat|) accumulate
// in the warp and enter the block reduction once.
// (<= 72 registers: 256 x 72 fit into the 19 K registers a resident contraction CTA leaves free on its SM)
template <int NPL>
__global__ void __launch_bounds__(kRowsPerBlock* kWarp) __maxnreg__(NPL <= 8 ? 72 : 168)
vat_finalize_kernel(const float* __restrict__ g, const float* __restrict__ d, const float* __restrict__ x,
                    float* __restrict__ r_adv, float* __restrict__ x_adv, float* __restrict__ d_hat,
                    int64_t n_rows, int row_len, float xi, float eps, float scale, int do_clamp,
                    int32_t* status_flag, float* __restrict__ dhat_abs_mean, float* __restrict__ workspace) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * kRowsPerBlock;
  int64_t row = (int64_t)blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  bool valid = row < n_rows;                               // warp-uniform
  RowStat st = {0u, 0.f};
  RowRegs<NPL> nd, nx, ng;
  if (valid) {
    load_row(nd, d + row * row_len, row_len, lane);
    load_row(nx, x + row * row_len, row_len, lane);
    load_row(ng, g + row * row_len, row_len, lane);
  }
  while (valid) {
    RowRegs<NPL> rd = nd, rx = nx, rg = ng;
    const int64_t off = row * row_len, next = row + stride;
    valid = next < n_rows;
    if (valid) {
      load_row(nd, d + next * row_len, row_len, lane);
      load_row(nx, x + next * row_len, row_len, lane);
      load_row(ng, g + next * row_len, row_len, lane);
    }
    const float n = sqrtf(row_sumsq(rd));
    const float rn = 1.f / n;
    // gd = xi * g * [0 <= x + xi*d/n <= 1]   (clamp passes the gradient on the closed interval)
    auto body = [&](auto fast) {
      constexpr bool kFast = decltype(fast)::value;
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        float gm = rg.v[i];
        if (do_clamp) {
          float s = rx.v[i] + xi * div_by<kFast>(rd.v[i], n, rn);
          gm = (s >= 0.f && s <= 1.f) ? gm : 0.f;
        }
        float gd = xi * gm;
        rg.v[i] = gd;
        dot = fmaf(gd, rd.v[i], dot);
      }
      dot = warp_sum(dot);
      const float c = dot / (n * n * n);
#pragma unroll
      for (int i = 0; i < NPL; ++i) rd.v[i] = (div_by<kFast>(rg.v[i], n, rn) - rd.v[i] * c) * scale;   // d.grad * scale
    };
    if (rcp_usable(rn)) body(std::true_type{}); else body(std::false_type{});
    const RowStat one = finalize_row(rd, rx, r_adv, x_adv, d_hat, off, row_len, lane, eps, do_clamp);
    st.bad |= one.bad;
    st.abs_sum += one.abs_sum;
    row = next;
  }
  finalize_block_stats(st, true, status_flag, dhat_abs_mean, workspace, (double)n_rows * row_len);
}

template <int NPL>
__global__ void __launch_bounds__(kRowsPerBlock* kWarp)
vat_direct_kernel(const float* __restrict__ d, const float* __restrict__ x, float* __restrict__ r_adv,
                  float* __restrict__ x_adv, float* __restrict__ d_hat, int64_t n_rows, int row_len, float eps,
                  int do_clamp, int32_t* status_flag, float* __restrict__ dhat_abs_mean, float* __restrict__ workspace) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  const bool valid = row < n_rows;                         // warp-uniform
  RowStat st = {0u, 0.f};
  if (valid) {
    const int64_t off = row * row_len;
    RowRegs<NPL> rd, rx;
    load_row(rd, d + off, row_len, lane);
    load_row(rx, x + off, row_len, lane);
    st = finalize_row(rd, rx, r_adv, x_adv, d_hat, off, row_len, lane, eps, do_clamp);
  }
  finalize_block_stats(st, valid, status_flag, dhat_abs_mean, workspace, (double)n_rows * row_len);
}

// ---- binwise=True flavour of _l2_normalize: d / (|d| + 1e-8), no row coupling (self_attention_VAT.py:242-243) ----
// Mathematically d/dd [d / (|d| + e)] = e / (|d| + e)^2 ~ 1e-8, but autograd forms it as the difference of two O(1)
// terms, go/b - go*((d/b)/b)*sgn(d) with b = |d| + e, which cancels to fp32 rounding noise: the reference's binwise
// direction IS that noise.  Parity therefore means executing the same IEEE op sequence (ATen's mul / div / abs
// backward formulas) with contraction disabled (__f*_rn intrinsics are never fused); with the same g the result is
// bit-identical to the reference's (tests: golden r_adv, torch.equal).
__device__ __forceinline__ float binwise_norm(float d) { return __fdiv_rn(d, __fadd_rn(fabsf(d), 1e-8f)); }

__global__ void __launch_bounds__(256)
vat_perturb_binwise_kernel(const float* __restrict__ x, const float* __restrict__ d, float* __restrict__ x_adv,
                           int64_t n, float xi, int do_clamp) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float s = __fadd_rn(__ldg(x + i), __fmul_rn(xi, binwise_norm(__ldg(d + i))));
    x_adv[i] = do_clamp ? clamp01(s) : s;
  }
}

// have_g = 1: d' = scale * d.grad (sequence above), then r_adv = eps * d' / (|d'| + 1e-8), x_adv, dhat.
// have_g = 0 (n_power == 0): d' = d.
__global__ void __launch_bounds__(256)
vat_finalize_binwise_kernel(const float* __restrict__ g, const float* __restrict__ d, const float* __restrict__ x,
                            float* __restrict__ r_adv, float* __restrict__ x_adv, float* __restrict__ d_hat,
                            int64_t n, float xi, float eps, float scale, int do_clamp, int have_g,
                            int32_t* status_flag) {
  unsigned bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float dv = __ldg(d + i), xv = __ldg(x + i);
    float dp = dv;
    if (have_g) {
      const float b = __fadd_rn(fabsf(dv), 1e-8f);
      const float q = __fdiv_rn(dv, b);
      float go = __ldg(g + i);
      if (do_clamp) {
        const float s = __fadd_rn(xv, __fmul_rn(xi, q));
        go = (s >= 0.f && s <= 1.f) ? go : 0.f;                         // clamp backward
      }
      const float gdn = __fmul_rn(go, xi);                              // mul backward
      const float grad_a = __fdiv_rn(gdn, b);                           // div backward, numerator
      const float grad_b = __fmul_rn(-gdn, __fdiv_rn(q, b));            // div backward, denominator
      const float sgn = (dv > 0.f) ? 1.f : ((dv < 0.f) ? -1.f : dv);    // torch.sgn: 0 -> 0, NaN -> NaN
      dp = __fmul_rn(__fadd_rn(grad_a, __fmul_rn(grad_b, sgn)), scale);
    }
    const float dh = binwise_norm(dp);
    const float r = __fmul_rn(eps, dh);
    bad |= (isnan(r) ? 1u : 0u) | (isinf(r) ? 2u : 0u);
    const float s2 = __fadd_rn(xv, r);
    r_adv[i] = r;
    x_adv[i] = do_clamp ? clamp01(s2) : s2;
    d_hat[i] = dh;
  }
  bad = __reduce_or_sync(kFull, bad);
  if (bad && (threadIdx.x & 31) == 0 && status_flag) atomicOr(status_flag, (int)bad);
}

// ---- V2: d divergence / d p -------------------------------------------------------------
// kind: RVB_DIV_BCE   F.binary_cross_entropy(p, y)                      (model/self_attention_VAT.py:182,200)
//       RVB_DIV_BKL   binary_kl_div(p, y): both clamped to [1e-4, 0.9999], F.kl_div(log [y, 1-y], [p, 1-p],
//                     'batchmean')                                       (model/self_attention_VAT.py:248-255)
//       RVB_DIV_MSE   F.mse_loss(p, y)                                   (model/onset_frame_VAT.py:232)
template <int kKind>
__device__ __forceinline__ float div_grad_one(float p, float y, float s) {
  if constexpr (kKind == RVB_DIV_BCE) {
    return (p - y) / max_nan((1.f - p) * p, 1e-12f) * s;                  // ATen's binary_cross_entropy_backward
  } else if constexpr (kKind == RVB_DIV_BKL) {
    // d/dq0 [q0 (log q0 - log p0) + q1 (log q1 - log p1)], q1 = 1 - q0; torch.clamp passes the gradient on the
    // closed interval only
    const float q0 = clamp_nan(p, 1e-4f, 0.9999f), p0 = clamp_nan(y, 1e-4f, 0.9999f);
    const float q1 = 1.f - q0, p1 = 1.f - p0;
    const float g = (logf(q0) - logf(p0)) - (logf(q1) - logf(p1));
    return (p >= 1e-4f && p <= 0.9999f) ? g * s : (p != p ? p : 0.f);    // NaN in -> NaN out, as autograd does
  } else {
    return 2.f * (p - y) * s;
  }
}

template <int kKind>
__global__ void __launch_bounds__(256)
div_grad_kernel(const float* __restrict__ p, const float* __restrict__ y, float* __restrict__ grad, int64_t n,
                double denom, const float* __restrict__ gscale_dev, float gscale, int vec_ok) {
  const float s = (gscale_dev ? __ldg(gscale_dev) : 1.f) * gscale / (float)denom;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    const float4* p4 = reinterpret_cast<const float4*>(p);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    float4* g4 = reinterpret_cast<float4*>(grad);
    for (int64_t j = i; j < n4; j += stride) {
      float4 a = __ldg(p4 + j), b = __ldg(y4 + j), o;
      o.x = div_grad_one<kKind>(a.x, b.x, s);
      o.y = div_grad_one<kKind>(a.y, b.y, s);
      o.z = div_grad_one<kKind>(a.z, b.z, s);
      o.w = div_grad_one<kKind>(a.w, b.w, s);
      g4[j] = o;
    }
    for (int64_t j = (n4 << 2) + i; j < n; j += stride) grad[j] = div_grad_one<kKind>(__ldg(p + j), __ldg(y + j), s);
  } else {
    for (int64_t j = i; j < n; j += stride) grad[j] = div_grad_one<kKind>(__ldg(p + j), __ldg(y + j), s);
  }
}

// ---- V4: divergence value (sum / denom), deterministic -----------------------------------
template <int kKind>
__device__ __forceinline__ float bce_one(float p, float y) {
  if constexpr (kKind == RVB_DIV_BCE) {
    // ATen: (y - 1) * max(log1p(-p), -100) - y * max(log(p), -100).  Logarithms via MUFU.LG2 (absolute error < 2^-21
    // outside [0.5, 2], 1 ulp inside): libm's logf + log1pf are ~80 instructions per element and made this 14 MB
    // reduction compute-bound (4.7 M warp instructions, profiles/r01f); the per-element error is unbiased and four
    // orders below the 1e-3 budget of the VAT loss.  log(0) = -inf and the -100 clamp behave as in ATen.
    // max_nan / clamp_nan: like ATen's std::max and torch.clamp, a NaN posterior gives a NaN loss (fmaxf would turn it
    // into a finite 100 and hide a diverged network from the logged VAT loss)
    return (y - 1.f) * max_nan(__logf(1.f - p), -100.f) - y * max_nan(__logf(p), -100.f);
  } else if constexpr (kKind == RVB_DIV_BKL) {
    // kl_div(input = log [p0, p1], target = [q0, q1]) pointwise: q log q - q input, summed over the pair
    const float q0 = clamp_nan(p, 1e-4f, 0.9999f), p0 = clamp_nan(y, 1e-4f, 0.9999f);
    const float q1 = 1.f - q0, p1 = 1.f - p0;
    return (q0 * logf(q0) - q0 * logf(p0)) + (q1 * logf(q1) - q1 * logf(p1));
  } else {
    const float e = p - y;
    return e * e;
  }
}

constexpr int kBceMaxBlocks = 1024;

template <int kKind>
__global__ void __launch_bounds__(256)
div_mean_kernel(const float* __restrict__ p, const float* __restrict__ y, int64_t n, double denom,
                float* __restrict__ loss, float* __restrict__ workspace, int vec_ok) {
  __shared__ float warp_part[8];
  __shared__ bool is_last;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    const float4* p4 = reinterpret_cast<const float4*>(p);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    // four float4 pairs per trip, all eight loads issued before the first logarithm: the kernel is latency-bound
    // (14 MB over 148 SMs), what counts is bytes in flight per thread
    constexpr int kU = 4;
    for (int64_t j0 = i; j0 < n4; j0 += stride * kU) {
      float4 a[kU], b[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int64_t j = j0 + u * stride;
        const bool ok = j < n4;
        a[u] = ok ? __ldg(p4 + j) : make_float4(1.f, 1.f, 1.f, 1.f);     // (p, y) = (1, 1) contributes exactly 0 to
        b[u] = ok ? __ldg(y4 + j) : make_float4(1.f, 1.f, 1.f, 1.f);     // every divergence
      }
#pragma unroll
      for (int u = 0; u < kU; ++u)
        if (j0 + u * stride < n4)
          acc += (bce_one<kKind>(a[u].x, b[u].x) + bce_one<kKind>(a[u].y, b[u].y)) +
                 (bce_one<kKind>(a[u].z, b[u].z) + bce_one<kKind>(a[u].w, b[u].w));
    }
    for (int64_t j = (n4 << 2) + i; j < n; j += stride) acc += bce_one<kKind>(__ldg(p + j), __ldg(y + j));
  } else {
    for (int64_t j = i; j < n; j += stride) acc += bce_one<kKind>(__ldg(p + j), __ldg(y + j));
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
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    __shared__ double dsum[8];
    if ((threadIdx.x & 31) == 0) dsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += dsum[w];
      *loss = (float)(tot / denom);
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

// Grid of the persistent row kernels: every warp gets at least one row, at most kPersistBlocksPerSM blocks per SM.
static unsigned persistent_grid(int64_t n_rows) {
  static int sms[kMaxDevices] = {};
  const int slot = device_slot();
  int n;
  {
    std::lock_guard<std::mutex> g(attr_mutex());
    if (sms[slot] == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms[slot], cudaDevAttrMultiProcessorCount, dev);
      if (sms[slot] <= 0) sms[slot] = 148;
    }
    n = sms[slot];
  }
  const int64_t need = (n_rows + kRowsPerBlock - 1) / kRowsPerBlock, cap = (int64_t)n * kPersistBlocksPerSM;
  return (unsigned)(need < cap ? need : cap);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace rvb

using namespace rvb;

extern "C" int rvb_vat_perturb(const float* x, const float* d, float* x_adv, int64_t n_rows, int row_len, float xi,
                               int do_clamp, rvb_stream_t stream) {
  RVB_REQUIRE(x && d && x_adv, "rvb_vat_perturb: null pointer");
  RVB_REQUIRE(n_rows >= 0 && row_len > 0, "rvb_vat_perturb: bad shape (%lld, %d)", (long long)n_rows, row_len);
  if (n_rows == 0) return RVB_OK;
  const unsigned grid = persistent_grid(n_rows);
  return dispatch_npl(row_len, [&](auto npl) {
    vat_perturb_kernel<decltype(npl)::value><<<grid, kRowsPerBlock * kWarp, 0, (cudaStream_t)stream>>>(
        x, d, x_adv, n_rows, row_len, xi, do_clamp);
    count_launch();
    return check_launch("vat_perturb_kernel");
  });
}

extern "C" int rvb_randn_like(float* out, int64_t n, uint64_t seed, uint64_t offset, uint32_t aten_threads,
                              uint64_t increment, uint64_t* dev_state, rvb_stream_t stream) {
  RVB_REQUIRE(out, "rvb_randn_like: null pointer");
  RVB_REQUIRE(n >= 0, "rvb_randn_like: negative size");
  RVB_REQUIRE(aten_threads > 0 && aten_threads % 256 == 0, "rvb_randn_like: aten_threads must be 256 * grid");
  RVB_REQUIRE((offset & 3) == 0 && (increment & 3) == 0, "rvb_randn_like: Philox offsets are multiples of 4");
  if (n == 0) return RVB_OK;
  DrawArgs a;
  a.seed = seed; a.offset = offset; a.dev_state = reinterpret_cast<unsigned long long*>(dev_state);
  a.increment = increment; a.tt = aten_threads;
  randn_like_kernel<4><<<aten_threads / 256, 256, 0, (cudaStream_t)stream>>>(out, n, a);
  count_launch();
  return check_launch("randn_like_kernel");
}

extern "C" int rvb_vat_perturb_draw(const float* x, float* d_out, float* x_adv, int64_t n_rows, int row_len, float xi,
                                    int do_clamp, uint64_t seed, uint64_t offset, uint32_t aten_threads,
                                    uint64_t increment, uint64_t* dev_state, rvb_stream_t stream) {
  RVB_REQUIRE(x && x_adv, "rvb_vat_perturb_draw: null pointer");
  RVB_REQUIRE(n_rows >= 0 && row_len > 0, "rvb_vat_perturb_draw: bad shape (%lld, %d)", (long long)n_rows, row_len);
  RVB_REQUIRE(aten_threads > 0 && aten_threads % 256 == 0, "rvb_vat_perturb_draw: aten_threads must be 256 * grid");
  RVB_REQUIRE((offset & 3) == 0 && (increment & 3) == 0, "rvb_vat_perturb_draw: Philox offsets are multiples of 4");
  RVB_REQUIRE(n_rows * (int64_t)row_len < (int64_t)aten_threads * 0x7fffffffll, "rvb_vat_perturb_draw: tensor too large");
  if (n_rows == 0) return RVB_OK;
  const unsigned grid = (unsigned)((n_rows + kRowsPerBlock - 1) / kRowsPerBlock);
  DrawArgs a;
  a.seed = seed; a.offset = offset; a.dev_state = reinterpret_cast<unsigned long long*>(dev_state);
  a.increment = increment; a.tt = aten_threads;
  return dispatch_npl(row_len, [&](auto npl) {
    vat_perturb_draw_kernel<decltype(npl)::value><<<grid, kRowsPerBlock * kWarp, 0, (cudaStream_t)stream>>>(
        x, d_out, x_adv, n_rows, row_len, xi, do_clamp, a);
    count_launch();
    return check_launch("vat_perturb_draw_kernel");
  });
}

static int launch_vat_finalize(const char* who, const float* g, const float* d, const float* x, float* r_adv,
                               float* x_adv, float* d_hat, int64_t n_rows, int row_len, float xi, float eps, float scale,
                               int do_clamp, int32_t* status_flag, float* dhat_abs_mean, float* workspace,
                               rvb_stream_t stream) {
  RVB_REQUIRE(d && x && r_adv && x_adv && (d_hat || workspace), "%s: null pointer", who);
  RVB_REQUIRE(n_rows >= 0 && row_len > 0, "%s: bad shape (%lld, %d)", who, (long long)n_rows, row_len);
  RVB_REQUIRE(n_rows < (int64_t)kRowsPerBlock * 0x7fffffff, "%s: too many rows", who);
  if (n_rows == 0) return RVB_OK;
  const unsigned grid = (unsigned)((n_rows + kRowsPerBlock - 1) / kRowsPerBlock);
  return dispatch_npl(row_len, [&](auto npl) {
    constexpr int N = decltype(npl)::value;
    if (g)
      vat_finalize_kernel<N><<<persistent_grid(n_rows), kRowsPerBlock * kWarp, 0, (cudaStream_t)stream>>>(
          g, d, x, r_adv, x_adv, d_hat, n_rows, row_len, xi, eps, scale, do_clamp, status_flag, dhat_abs_mean, workspace);
    else
      vat_direct_kernel<N><<<grid, kRowsPerBlock * kWarp, 0, (cudaStream_t)stream>>>(
          d, x, r_adv, x_adv, d_hat, n_rows, row_len, eps, do_clamp, status_flag, dhat_abs_mean, workspace);
    count_launch();
    return check_launch(g ? "vat_finalize_kernel" : "vat_direct_kernel");
  });
}

extern "C" int rvb_vat_finalize(const float* g, const float* d, const float* x, float* r_adv, float* x_adv,
                                float* d_hat, int64_t n_rows, int row_len, float xi, float eps, float scale,
                                int do_clamp, int32_t* status_flag, rvb_stream_t stream) {
  RVB_REQUIRE(g, "rvb_vat_finalize: null pointer");
  return launch_vat_finalize("rvb_vat_finalize", g, d, x, r_adv, x_adv, d_hat, n_rows, row_len, xi, eps, scale, do_clamp,
                             status_flag, nullptr, nullptr, stream);
}

extern "C" int rvb_vat_direct(const float* d, const float* x, float* r_adv, float* x_adv, float* d_hat,
                              int64_t n_rows, int row_len, float eps, int do_clamp, int32_t* status_flag,
                              rvb_stream_t stream) {
  return launch_vat_finalize("rvb_vat_direct", nullptr, d, x, r_adv, x_adv, d_hat, n_rows, row_len, 0.f, eps, 1.f,
                             do_clamp, status_flag, nullptr, nullptr, stream);
}

extern "C" int64_t rvb_vat_stats_workspace_bytes(int64_t n_rows) {
  const int64_t blocks = (n_rows + kRowsPerBlock - 1) / kRowsPerBlock;
  return (2 * blocks + 1) * 4;
}

extern "C" int rvb_vat_finalize_stats(const float* g, const float* d, const float* x, float* r_adv, float* x_adv,
                                      float* d_hat, int64_t n_rows, int row_len, float xi, float eps, float scale,
                                      int do_clamp, int32_t* status_flag, float* dhat_abs_mean, void* workspace,
                                      int64_t workspace_bytes, rvb_stream_t stream) {
  RVB_REQUIRE(status_flag && dhat_abs_mean && workspace, "rvb_vat_finalize_stats: null pointer");
  RVB_REQUIRE(n_rows > 0, "rvb_vat_finalize_stats: empty input (the reference's mean of nothing is NaN)");
  RVB_REQUIRE(workspace_bytes >= rvb_vat_stats_workspace_bytes(n_rows),
              "rvb_vat_finalize_stats: workspace of %lld bytes, %lld needed", (long long)workspace_bytes,
              (long long)rvb_vat_stats_workspace_bytes(n_rows));
  return launch_vat_finalize("rvb_vat_finalize_stats", g, d, x, r_adv, x_adv, d_hat, n_rows, row_len, xi, eps, scale,
                             do_clamp, status_flag, dhat_abs_mean, static_cast<float*>(workspace), stream);
}

static unsigned flat_grid(int64_t n, int per_thread) {
  int64_t blocks = (n + 256LL * per_thread - 1) / (256LL * per_thread);
  // a few waves of 148 SMs x 8 resident blocks is plenty for a streaming kernel
  const int64_t cap = 148 * 16;
  return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

template <typename F>
static int dispatch_kind(int kind, const char* who, F&& f) {
  switch (kind) {
    case RVB_DIV_BCE: return f(std::integral_constant<int, RVB_DIV_BCE>{});
    case RVB_DIV_BKL: return f(std::integral_constant<int, RVB_DIV_BKL>{});
    case RVB_DIV_MSE: return f(std::integral_constant<int, RVB_DIV_MSE>{});
    default: set_error("%s: unknown divergence kind %d", who, kind); return RVB_ERR_ARG;
  }
}

extern "C" int rvb_div_grad(int kind, const float* p, const float* y, float* grad, int64_t n, double denom,
                            const float* gscale_dev, float gscale, rvb_stream_t stream) {
  RVB_REQUIRE(p && y && grad, "rvb_div_grad: null pointer");
  RVB_REQUIRE(n >= 0 && denom > 0, "rvb_div_grad: bad size");
  if (n == 0) return RVB_OK;
  const int vec = aligned16(p) && aligned16(y) && aligned16(grad);
  return dispatch_kind(kind, "rvb_div_grad", [&](auto k) {
    div_grad_kernel<decltype(k)::value><<<flat_grid(n, 8), 256, 0, (cudaStream_t)stream>>>(p, y, grad, n, denom,
                                                                                            gscale_dev, gscale, vec);
    count_launch();
    return check_launch("div_grad_kernel");
  });
}

extern "C" int rvb_div_mean(int kind, const float* p, const float* y, int64_t n, double denom, float* loss,
                            float* workspace, rvb_stream_t stream) {
  RVB_REQUIRE(p && y && loss && workspace, "rvb_div_mean: null pointer");
  RVB_REQUIRE(n > 0 && denom > 0, "rvb_div_mean: empty input (the reference returns NaN for an empty mean)");
  // one wave: at most two 256-thread blocks per SM, every thread looping over trips of four float4 pairs -- the last-
  // block reduction then reads ~300 partials instead of ~900, and the grid has no tail
  unsigned grid = flat_grid(n, 16);
  const unsigned cap = 2 * persistent_grid(1ll << 40) / kPersistBlocksPerSM;
  if (grid > cap) grid = cap;
  if (grid > (unsigned)kBceMaxBlocks) grid = kBceMaxBlocks;
  const int vec = aligned16(p) && aligned16(y);
  return dispatch_kind(kind, "rvb_div_mean", [&](auto k) {
    div_mean_kernel<decltype(k)::value><<<grid, 256, 0, (cudaStream_t)stream>>>(p, y, n, denom, loss, workspace, vec);
    count_launch();
    return check_launch("div_mean_kernel");
  });
}

extern "C" int rvb_bce_grad(const float* p, const float* y, float* grad, int64_t n, const float* gscale_dev,
                            float gscale, rvb_stream_t stream) {
  return rvb_div_grad(RVB_DIV_BCE, p, y, grad, n, (double)(n > 0 ? n : 1), gscale_dev, gscale, stream);
}

extern "C" int rvb_bce_mean(const float* p, const float* y, int64_t n, float* loss, float* workspace,
                            rvb_stream_t stream) {
  RVB_REQUIRE(n > 0, "rvb_bce_mean: empty input (the reference returns NaN for an empty mean)");
  return rvb_div_mean(RVB_DIV_BCE, p, y, n, (double)n, loss, workspace, stream);
}

extern "C" int rvb_vat_perturb_binwise(const float* x, const float* d, float* x_adv, int64_t n, float xi, int do_clamp,
                                       rvb_stream_t stream) {
  RVB_REQUIRE(x && d && x_adv, "rvb_vat_perturb_binwise: null pointer");
  RVB_REQUIRE(n >= 0, "rvb_vat_perturb_binwise: negative size");
  if (n == 0) return RVB_OK;
  vat_perturb_binwise_kernel<<<flat_grid(n, 4), 256, 0, (cudaStream_t)stream>>>(x, d, x_adv, n, xi, do_clamp);
  count_launch();
  return check_launch("vat_perturb_binwise_kernel");
}

extern "C" int rvb_vat_finalize_binwise(const float* g, const float* d, const float* x, float* r_adv, float* x_adv,
                                        float* d_hat, int64_t n, float xi, float eps, float scale, int do_clamp,
                                        int32_t* status_flag, rvb_stream_t stream) {
  RVB_REQUIRE(d && x && r_adv && x_adv && d_hat, "rvb_vat_finalize_binwise: null pointer");
  RVB_REQUIRE(n >= 0, "rvb_vat_finalize_binwise: negative size");
  if (n == 0) return RVB_OK;
  vat_finalize_binwise_kernel<<<flat_grid(n, 4), 256, 0, (cudaStream_t)stream>>>(g, d, x, r_adv, x_adv, d_hat, n, xi, eps,
                                                                               scale, do_clamp, g != nullptr, status_flag);
  count_launch();
  return check_launch("vat_finalize_binwise_kernel");
}
