// Train-mode BatchNorm2d of the caller's U-Net (SURVEY.md 8f, row f5 -- consumer side of the hot path).
//
// model/self_attention_VAT.py:848-850, :865-869 put an nn.BatchNorm2d(C = 16 .. 128) behind every convolution of the
// two U-Nets: 30 modules, three forward and two backward passes per training iteration.  cuDNN's spatial kernels for
// this layout (bn_fw_tr_1C11_kernel_NCHW / bn_bw_1C11_kernel_new) launch ONE block per channel -- 16 blocks on 148
// SMs for the (8, 16, 640, 229) tensors -- and take 52 % of the iteration on a B200 (profiles/r02_train_step.md).  The
// op is two reductions and an elementwise pass over 75 MB: an HBM problem.  Here every channel is cut into `splits`
// slices so that ~4 blocks per SM run, partial sums go through a small float64 workspace, and the pass that applies
// the result re-derives the channel statistics from those partials in its prologue -- two launches per direction.
//
//   forward  (training)  rvb_bn_reduce(x)              partial sums of (x - K), (x - K)^2, K = first element of the channel
//                        rvb_bn_forward                mean, biased var, y = (x - mean) invstd gamma + beta, running stats
//   forward  (eval)      rvb_bn_apply                  y = (x - mean) invstd gamma + beta with the given statistics
//   backward             rvb_bn_reduce(x, dy, mean)    partial sums of dy, dy (x - mean)
//                        rvb_bn_backward               dx, dgamma, dbeta   (ATen: native_batch_norm_backward)
#include "rvb_common.cuh"

namespace rvb {

extern void count_launch();

constexpr int kBnThreads = 256;
constexpr int kBnMaxSplits = 64;

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// Block-wide sum of two doubles; the result is valid in thread 0.
__device__ __forceinline__ void block_sum2(double& a, double& b) {
  __shared__ double s_part[kBnThreads / kWarp][2];
  a = warp_sum_f64(a);
  b = warp_sum_f64(b);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_part[warp][0] = a; s_part[warp][1] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = 0.0; b = 0.0;
#pragma unroll
    for (int w = 0; w < kBnThreads / kWarp; ++w) { a += s_part[w][0]; b += s_part[w][1]; }
  }
}

// grid (splits, C).  Block (s, ch) covers elements [s chunk, min(hw, (s + 1) chunk)) of every sample's channel ch.
template <bool kBackward>
__global__ void __launch_bounds__(kBnThreads)
bn_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mean, int n, int c,
                 int64_t hw, int64_t chunk, int vec, double* __restrict__ partials) {
  const int ch = blockIdx.y, s = blockIdx.x;
  const int64_t j0 = (int64_t)s * chunk;
  const int64_t j1 = j0 + chunk < hw ? j0 + chunk : hw;
  const float K = kBackward ? __ldg(mean + ch) : __ldg(x + (int64_t)ch * hw);
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < n; ++i) {
    const int64_t base = ((int64_t)i * c + ch) * hw;
    if (vec) {
      auto add = [&](const float4& v, const float4& g) {
        const float d[4] = {v.x - K, v.y - K, v.z - K, v.w - K};
        if constexpr (kBackward) {
          const float gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) { a[q] += gg[q]; b[q] = fmaf(gg[q], d[q], b[q]); }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) { a[q] += d[q]; b[q] = fmaf(d[q], d[q], b[q]); }
        }
      };
      constexpr int64_t kStep = 4 * kBnThreads;
      int64_t j = j0 + 4 * (int64_t)threadIdx.x;
      for (; j < j1; j += 4 * kStep) {                       // up to four independent 16-byte loads (eight in the backward)
        float4 v[4], g[4];                                   // in flight; a slice is only ~4 trips long per sample
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (j + u * kStep < j1) {
            v[u] = __ldg(reinterpret_cast<const float4*>(x + base + j + u * kStep));
            if constexpr (kBackward) g[u] = __ldg(reinterpret_cast<const float4*>(dy + base + j + u * kStep));
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j + u * kStep < j1) add(v[u], g[u]);
      }
    } else {
      for (int64_t j = j0 + threadIdx.x; j < j1; j += kBnThreads) {
        const float d = __ldg(x + base + j) - K;
        const int q = (int)((j - j0) / kBnThreads) & 3;
        if constexpr (kBackward) {
          const float g = __ldg(dy + base + j);
          a[q] += g; b[q] = fmaf(g, d, b[q]);
        } else {
          a[q] += d; b[q] = fmaf(d, d, b[q]);
        }
      }
    }
  }
  double sa = ((double)a[0] + (double)a[1]) + ((double)a[2] + (double)a[3]);
  double sb = ((double)b[0] + (double)b[1]) + ((double)b[2] + (double)b[3]);
  block_sum2(sa, sb);
  if (threadIdx.x == 0) {
    partials[2 * ((int64_t)ch * gridDim.x + s)] = sa;
    partials[2 * ((int64_t)ch * gridDim.x + s) + 1] = sb;
  }
}

// Sum of this channel's partials, by warp 0; valid in lane 0.
__device__ __forceinline__ void channel_sums(const double* __restrict__ partials, int ch, int splits, double& s1, double& s2) {
  const int lane = threadIdx.x & 31;
  s1 = 0.0; s2 = 0.0;
  for (int i = lane; i < splits; i += kWarp) {
    s1 += partials[2 * ((int64_t)ch * splits + i)];
    s2 += partials[2 * ((int64_t)ch * splits + i) + 1];
  }
  s1 = warp_sum_f64(s1);
  s2 = warp_sum_f64(s2);
}

// y = (x - mean) * scale + shift over this block's slice
__device__ __forceinline__ void apply_slice(const float* __restrict__ x, float* __restrict__ y, int n, int c, int ch,
                                            int64_t hw, int64_t j0, int64_t j1, int vec, float mean, float scale, float shift) {
  for (int i = 0; i < n; ++i) {
    const int64_t base = ((int64_t)i * c + ch) * hw;
    if (vec) {
      auto one = [&](const float4& v) {
        float4 o;
        o.x = fmaf(v.x - mean, scale, shift); o.y = fmaf(v.y - mean, scale, shift);
        o.z = fmaf(v.z - mean, scale, shift); o.w = fmaf(v.w - mean, scale, shift);
        return o;
      };
      constexpr int64_t kStep = 4 * kBnThreads;
      int64_t j = j0 + 4 * (int64_t)threadIdx.x;
      for (; j < j1; j += 4 * kStep) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j + u * kStep < j1) v[u] = __ldg(reinterpret_cast<const float4*>(x + base + j + u * kStep));
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j + u * kStep < j1) *reinterpret_cast<float4*>(y + base + j + u * kStep) = one(v[u]);
      }
    } else {
      for (int64_t j = j0 + threadIdx.x; j < j1; j += kBnThreads) y[base + j] = fmaf(__ldg(x + base + j) - mean, scale, shift);
    }
  }
}

__global__ void __launch_bounds__(kBnThreads)
bn_forward_kernel(const float* __restrict__ x, int n, int c, int64_t hw, int64_t chunk, int vec,
                  const double* __restrict__ partials, const float* __restrict__ gamma, const float* __restrict__ beta,
                  float eps, float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                  float* __restrict__ save_mean, float* __restrict__ save_invstd, float* __restrict__ y) {
  const int ch = blockIdx.y, s = blockIdx.x, splits = gridDim.x;
  __shared__ float s_stat[3];
  if (threadIdx.x < kWarp) {
    double s1, s2;
    channel_sums(partials, ch, splits, s1, s2);
    if (threadIdx.x == 0) {
      const double cnt = (double)n * (double)hw;
      const double K = (double)__ldg(x + (int64_t)ch * hw);
      const double m_sh = s1 / cnt;
      double var = s2 / cnt - m_sh * m_sh;
      if (var < 0.0) var = 0.0;
      const float mean = (float)(K + m_sh);
      const float invstd = (float)(1.0 / sqrt(var + (double)eps));
      const float g = gamma ? __ldg(gamma + ch) : 1.f;
      s_stat[0] = mean;
      s_stat[1] = invstd * g;
      s_stat[2] = beta ? __ldg(beta + ch) : 0.f;
      if (s == 0) {
        save_mean[ch] = mean;
        save_invstd[ch] = invstd;
        if (running_mean) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * mean;
        if (running_var) {
          const float unbiased = (float)(var * (cnt / (cnt - 1.0)));
          running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * unbiased;
        }
      }
    }
  }
  __syncthreads();
  const int64_t j0 = (int64_t)s * chunk;
  const int64_t j1 = j0 + chunk < hw ? j0 + chunk : hw;
  apply_slice(x, y, n, c, ch, hw, j0, j1, vec, s_stat[0], s_stat[1], s_stat[2]);
}

__global__ void __launch_bounds__(kBnThreads)
bn_apply_kernel(const float* __restrict__ x, int n, int c, int64_t hw, int64_t chunk, int vec, const float* __restrict__ mean,
                const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                float* __restrict__ y) {
  const int ch = blockIdx.y, s = blockIdx.x;
  const float g = gamma ? __ldg(gamma + ch) : 1.f;
  const int64_t j0 = (int64_t)s * chunk;
  const int64_t j1 = j0 + chunk < hw ? j0 + chunk : hw;
  apply_slice(x, y, n, c, ch, hw, j0, j1, vec, __ldg(mean + ch), __ldg(invstd + ch) * g, beta ? __ldg(beta + ch) : 0.f);
}

__global__ void __launch_bounds__(kBnThreads)
bn_backward_kernel(const float* __restrict__ x, const float* __restrict__ dy, int n, int c, int64_t hw, int64_t chunk, int vec,
                   const double* __restrict__ partials, const float* __restrict__ gamma, const float* __restrict__ mean,
                   const float* __restrict__ invstd, int training, float* __restrict__ dx, float* __restrict__ dgamma,
                   float* __restrict__ dbeta) {
  const int ch = blockIdx.y, s = blockIdx.x, splits = gridDim.x;
  __shared__ float s_k[3];
  const float mu = __ldg(mean + ch), is = __ldg(invstd + ch);
  if (threadIdx.x < kWarp) {
    double sdy, sdyx;
    channel_sums(partials, ch, splits, sdy, sdyx);
    if (threadIdx.x == 0) {
      const double cnt = (double)n * (double)hw;
      const float g = gamma ? __ldg(gamma + ch) : 1.f;
      s_k[0] = training ? (float)(sdy / cnt) : 0.f;
      s_k[1] = training ? (float)(sdyx * (double)is * (double)is / cnt) : 0.f;
      s_k[2] = is * g;
      if (s == 0) {
        if (dgamma) dgamma[ch] = (float)(sdyx * (double)is);
        if (dbeta) dbeta[ch] = (float)sdy;
      }
    }
  }
  __syncthreads();
  if (dx == nullptr) return;
  const float k1 = s_k[0], k2 = s_k[1], sc = s_k[2];
  const int64_t j0 = (int64_t)s * chunk;
  const int64_t j1 = j0 + chunk < hw ? j0 + chunk : hw;
  for (int i = 0; i < n; ++i) {
    const int64_t base = ((int64_t)i * c + ch) * hw;
    if (vec) {
      auto one = [&](const float4& v, const float4& g) {
        float4 o;
        o.x = (g.x - k1 - (v.x - mu) * k2) * sc; o.y = (g.y - k1 - (v.y - mu) * k2) * sc;
        o.z = (g.z - k1 - (v.z - mu) * k2) * sc; o.w = (g.w - k1 - (v.w - mu) * k2) * sc;
        return o;
      };
      constexpr int64_t kStep = 4 * kBnThreads;
      int64_t j = j0 + 4 * (int64_t)threadIdx.x;
      for (; j < j1; j += 4 * kStep) {
        float4 v[4], g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (j + u * kStep < j1) {
            v[u] = __ldg(reinterpret_cast<const float4*>(x + base + j + u * kStep));
            g[u] = __ldg(reinterpret_cast<const float4*>(dy + base + j + u * kStep));
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j + u * kStep < j1) *reinterpret_cast<float4*>(dx + base + j + u * kStep) = one(v[u], g[u]);
      }
    } else {
      for (int64_t j = j0 + threadIdx.x; j < j1; j += kBnThreads)
        dx[base + j] = (__ldg(dy + base + j) - k1 - (__ldg(x + base + j) - mu) * k2) * sc;
    }
  }
}

// ---------------------------------------------------------------- channels_last (NHWC) flavour
// x[p][c], p = pixel (n, h, w), c contiguous.  A block owns a range of pixels and ALL channels; a thread owns one group of
// four channels (float4) for the whole kernel -- blockDim is a multiple of C / 4 -- so its sums, and later its scale /
// shift, stay in registers.  Partials: [block][C][2] doubles; the LAST block to finish (ticket) adds them per channel
// and writes the channel statistics, so the applying kernel starts from 2-3 floats per channel.
struct BnNhwcWs {                       // workspace layout (doubles first); see rvb_bn_nhwc_workspace_bytes
  double* partials;                     // [kBnNhwcMaxBlocks][C][2]
  float* coef;                          // [3][C]: backward k1, k2, invstd * gamma
  unsigned int* ticket;                 // zero before the first launch; the last block resets it
};
constexpr int kBnNhwcMaxBlocks = 1024;

__device__ __forceinline__ BnNhwcWs nhwc_ws(void* ws, int c) {
  BnNhwcWs w;
  w.partials = reinterpret_cast<double*>(ws);
  w.coef = reinterpret_cast<float*>(w.partials + (size_t)kBnNhwcMaxBlocks * c * 2);
  w.ticket = reinterpret_cast<unsigned int*>(w.coef + 3 * (size_t)c);
  return w;
}

template <bool kBackward>
__global__ void __launch_bounds__(kBnThreads)
bn_nhwc_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy, int64_t pixels, int c, int64_t per_block,
                      void* __restrict__ ws_raw,
                      // forward finalisation
                      float eps, float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                      float* __restrict__ save_mean, float* __restrict__ save_invstd,
                      // backward finalisation
                      const float* __restrict__ gamma, int training, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ double s_red[];                        // [8][blockDim]
  const BnNhwcWs ws = nhwc_ws(ws_raw, c);
  const int cg = c >> 2;
  const int g = threadIdx.x % cg, r = threadIdx.x / cg, rows = blockDim.x / cg;
  const int64_t p0 = (int64_t)blockIdx.x * per_block;
  const int64_t p1 = p0 + per_block < pixels ? p0 + per_block : pixels;
  // shift: forward K = pixel 0 of the tensor; backward K = the saved mean
  const float4 K = kBackward ? __ldg(reinterpret_cast<const float4*>(save_mean) + g) : __ldg(reinterpret_cast<const float4*>(x) + g);
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  auto add = [&](const float4& v, const float4& gy) {
    const float d[4] = {v.x - K.x, v.y - K.y, v.z - K.z, v.w - K.w};
    if constexpr (kBackward) {
      const float gg[4] = {gy.x, gy.y, gy.z, gy.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) { a[q] += gg[q]; b[q] = fmaf(gg[q], d[q], b[q]); }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) { a[q] += d[q]; b[q] = fmaf(d[q], d[q], b[q]); }
    }
  };
  int64_t p = p0 + r;
  for (; p + 3 * rows < p1; p += 4 * rows) {               // four pixels per trip: independent loads in flight
    float4 v[4], gy[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u] = __ldg(reinterpret_cast<const float4*>(x + (p + u * rows) * c) + g);
      if constexpr (kBackward) gy[u] = __ldg(reinterpret_cast<const float4*>(dy + (p + u * rows) * c) + g);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) add(v[u], gy[u]);
  }
  for (; p < p1; p += rows) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + p * c) + g);
    float4 gy = v;
    if constexpr (kBackward) gy = __ldg(reinterpret_cast<const float4*>(dy + p * c) + g);
    add(v, gy);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    s_red[(2 * q) * blockDim.x + threadIdx.x] = (double)a[q];
    s_red[(2 * q + 1) * blockDim.x + threadIdx.x] = (double)b[q];
  }
  __syncthreads();
  if (r == 0) {                                            // one thread per channel group adds the rows
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double sa = 0.0, sb = 0.0;
      for (int rr = 0; rr < rows; ++rr) {
        sa += s_red[(2 * q) * blockDim.x + rr * cg + g];
        sb += s_red[(2 * q + 1) * blockDim.x + rr * cg + g];
      }
      const int ch = 4 * g + q;
      ws.partials[2 * ((int64_t)blockIdx.x * c + ch)] = sa;
      ws.partials[2 * ((int64_t)blockIdx.x * c + ch) + 1] = sb;
    }
  }
  // ---- last block: per-channel totals and statistics
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ws.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const double cnt = (double)pixels;
  // every thread adds a share of the per-block partials: thread t takes channel t % c' of a window of c' = min(c,
  // blockDim) channels and every nparts-th block; the shares meet in shared memory
  const int cw = c < (int)blockDim.x ? c : (int)blockDim.x;
  const int nparts = (int)blockDim.x / cw;
  for (int c0 = 0; c0 < c; c0 += cw) {
    const int chl = threadIdx.x % cw, part = threadIdx.x / cw;
    double t1 = 0.0, t2 = 0.0;
    if (part < nparts && c0 + chl < c) {
      for (unsigned blk = part; blk < gridDim.x; blk += nparts) {
        t1 += __ldcg(ws.partials + 2 * ((int64_t)blk * c + c0 + chl));
        t2 += __ldcg(ws.partials + 2 * ((int64_t)blk * c + c0 + chl) + 1);
      }
    }
    __syncthreads();
    s_red[threadIdx.x] = t1;
    s_red[blockDim.x + threadIdx.x] = t2;
    __syncthreads();
    if (threadIdx.x < cw && c0 + threadIdx.x < c) {
      const int ch = c0 + threadIdx.x;
      double s1 = 0.0, s2 = 0.0;
      for (int pp = 0; pp < nparts; ++pp) {
        s1 += s_red[pp * cw + threadIdx.x];
        s2 += s_red[blockDim.x + pp * cw + threadIdx.x];
      }
    if constexpr (kBackward) {
      const float is = save_invstd[ch];
      const float gm = gamma ? gamma[ch] : 1.f;
      ws.coef[ch] = training ? (float)(s1 / cnt) : 0.f;
      ws.coef[c + ch] = training ? (float)(s2 * (double)is * (double)is / cnt) : 0.f;
      ws.coef[2 * c + ch] = is * gm;
      if (dgamma) dgamma[ch] = (float)(s2 * (double)is);
      if (dbeta) dbeta[ch] = (float)s1;
    } else {
      const double Kc = (double)x[ch];
      const double m_sh = s1 / cnt;
      double var = s2 / cnt - m_sh * m_sh;
      if (var < 0.0) var = 0.0;
      const float mean = (float)(Kc + m_sh);
      save_mean[ch] = mean;
      save_invstd[ch] = (float)(1.0 / sqrt(var + (double)eps));
      if (running_mean) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * mean;
      if (running_var) running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)(var * (cnt / (cnt - 1.0)));
    }
    }
  }
  if (threadIdx.x == 0) *ws.ticket = 0u;                    // ready for the next launch (stream order)
}

// kMode 0: y = (x - mean) invstd gamma + beta;  1: dx = (dy - k1 - (x - mean) k2) * (invstd gamma) with coef from the workspace
template <int kMode>
__global__ void __launch_bounds__(kBnThreads)
bn_nhwc_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, int64_t pixels, int c, int64_t per_block,
                     const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const void* __restrict__ ws_raw, float* __restrict__ out) {
  const int cg = c >> 2;
  const int g = threadIdx.x % cg, r = threadIdx.x / cg, rows = blockDim.x / cg;
  const int64_t p0 = (int64_t)blockIdx.x * per_block;
  const int64_t p1 = p0 + per_block < pixels ? p0 + per_block : pixels;
  const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + g);
  float4 sc, sh, k2 = make_float4(0.f, 0.f, 0.f, 0.f);
  if constexpr (kMode == 0) {
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd) + g);
    const float4 gm = gamma ? __ldg(reinterpret_cast<const float4*>(gamma) + g) : make_float4(1.f, 1.f, 1.f, 1.f);
    sc = make_float4(is.x * gm.x, is.y * gm.y, is.z * gm.z, is.w * gm.w);
    sh = beta ? __ldg(reinterpret_cast<const float4*>(beta) + g) : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    const float* coef = nhwc_ws(const_cast<void*>(ws_raw), c).coef;
    sh = __ldcg(reinterpret_cast<const float4*>(coef) + g);             // k1
    k2 = __ldcg(reinterpret_cast<const float4*>(coef + c) + g);
    sc = __ldcg(reinterpret_cast<const float4*>(coef + 2 * c) + g);
  }
  auto one = [&](const float4& v, const float4& gy) {
    float4 o;
    if constexpr (kMode == 0) {
      o.x = fmaf(v.x - mu.x, sc.x, sh.x); o.y = fmaf(v.y - mu.y, sc.y, sh.y);
      o.z = fmaf(v.z - mu.z, sc.z, sh.z); o.w = fmaf(v.w - mu.w, sc.w, sh.w);
    } else {
      o.x = (gy.x - sh.x - (v.x - mu.x) * k2.x) * sc.x; o.y = (gy.y - sh.y - (v.y - mu.y) * k2.y) * sc.y;
      o.z = (gy.z - sh.z - (v.z - mu.z) * k2.z) * sc.z; o.w = (gy.w - sh.w - (v.w - mu.w) * k2.w) * sc.w;
    }
    return o;
  };
  int64_t p = p0 + r;
  for (; p + 3 * rows < p1; p += 4 * rows) {               // four pixels per trip
    float4 v[4], gy[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u] = __ldg(reinterpret_cast<const float4*>(x + (p + u * rows) * c) + g);
      if constexpr (kMode == 1) gy[u] = __ldg(reinterpret_cast<const float4*>(dy + (p + u * rows) * c) + g);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) *(reinterpret_cast<float4*>(out + (p + u * rows) * c) + g) = one(v[u], gy[u]);
  }
  for (; p < p1; p += rows) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + p * c) + g);
    float4 gy = v;
    if constexpr (kMode == 1) gy = __ldg(reinterpret_cast<const float4*>(dy + p * c) + g);
    *(reinterpret_cast<float4*>(out + p * c) + g) = one(v, gy);
  }
}

struct NhwcPlan { int threads, blocks; int64_t per_block; };

static int nhwc_plan(const char* who, int64_t pixels, int c, NhwcPlan* plan) {
  RVB_REQUIRE(pixels > 0 && c > 0 && c % 4 == 0 && c <= 4 * kBnThreads, "%s: the channels_last kernels need c %% 4 == 0 and c <= %d (got %d)",
              who, 4 * kBnThreads, c);
  const int cg = c / 4;
  plan->threads = kBnThreads / cg * cg;
  const int rows = plan->threads / cg;
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t blocks = 4 * (int64_t)sms;
  const int64_t most = (pixels + 8 * rows - 1) / (8 * rows);             // at least eight loop trips per thread
  if (blocks > most) blocks = most;
  if (blocks > kBnNhwcMaxBlocks) blocks = kBnNhwcMaxBlocks;
  if (blocks < 1) blocks = 1;
  plan->per_block = (pixels + blocks - 1) / blocks;
  plan->blocks = (int)((pixels + plan->per_block - 1) / plan->per_block);
  return RVB_OK;
}

static bool aligned16(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int bn_geometry(const char* who, int n, int c, int64_t hw, int splits, int64_t* chunk) {
  RVB_REQUIRE(n > 0 && c > 0 && c <= 65535 && hw > 0, "%s: bad shape (n %d, c %d, hw %lld)", who, n, c, (long long)hw);
  RVB_REQUIRE(splits >= 1 && splits <= kBnMaxSplits, "%s: splits %d outside [1, %d]", who, splits, kBnMaxSplits);
  const int64_t per = (hw + splits - 1) / splits;
  *chunk = (per + 3) / 4 * 4;
  RVB_REQUIRE((int64_t)(splits - 1) * *chunk < hw, "%s: %d splits leave an empty slice for hw %lld (use rvb_bn_splits)", who,
              splits, (long long)hw);
  return RVB_OK;
}

}  // namespace rvb

using namespace rvb;

extern "C" int rvb_bn_splits(int n, int c, int64_t hw) {
  if (n <= 0 || c <= 0 || hw <= 0) return 1;
  static int sm_count[kMaxDevices] = {};
  const int slot = device_slot();
  int sms = sm_count[slot];                                      // benign race: every writer stores the same value
  if (sms == 0) {
    int dev = 0;
    sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    sm_count[slot] = sms;
  }
  int64_t want = (4 * (int64_t)sms + c - 1) / c;                 // ~4 blocks per SM
  const int64_t most = hw / 1024 > 1 ? hw / 1024 : 1;            // slices of at least 1 024 elements per sample
  if (want > most) want = most;
  if (want > kBnMaxSplits) want = kBnMaxSplits;
  if (want < 1) want = 1;
  // no empty last slice: chunk = roundup4(ceil(hw / splits)) must leave (splits - 1) * chunk < hw
  while (want > 1) {
    const int64_t chunk = ((hw + want - 1) / want + 3) / 4 * 4;
    if ((want - 1) * chunk < hw) break;
    --want;
  }
  return (int)want;
}

extern "C" int rvb_bn_reduce(const float* x, const float* dy, const float* mean, int n, int c, int64_t hw, int splits,
                             double* partials, rvb_stream_t stream) {
  const char* who = "rvb_bn_reduce";
  RVB_REQUIRE(x && partials, "%s: null pointer", who);
  RVB_REQUIRE((dy == nullptr) == (mean == nullptr), "%s: dy and mean go together (backward sums)", who);
  int64_t chunk;
  int rc = bn_geometry(who, n, c, hw, splits, &chunk);
  if (rc != RVB_OK) return rc;
  const int vec = (hw % 4 == 0) && aligned16(x) && aligned16(dy);
  const dim3 grid(splits, c);
  if (dy)
    bn_reduce_kernel<true><<<grid, kBnThreads, 0, (cudaStream_t)stream>>>(x, dy, mean, n, c, hw, chunk, vec, partials);
  else
    bn_reduce_kernel<false><<<grid, kBnThreads, 0, (cudaStream_t)stream>>>(x, nullptr, nullptr, n, c, hw, chunk, vec, partials);
  count_launch();
  return check_launch("bn_reduce_kernel");
}

extern "C" int rvb_bn_forward(const float* x, int n, int c, int64_t hw, int splits, const double* partials,
                              const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                              float* running_var, float* save_mean, float* save_invstd, float* y, rvb_stream_t stream) {
  const char* who = "rvb_bn_forward";
  RVB_REQUIRE(x && partials && save_mean && save_invstd && y, "%s: null pointer", who);
  RVB_REQUIRE((int64_t)n * hw > 1, "%s: more than one value per channel is needed in training mode", who);
  int64_t chunk;
  int rc = bn_geometry(who, n, c, hw, splits, &chunk);
  if (rc != RVB_OK) return rc;
  const int vec = (hw % 4 == 0) && aligned16(x) && aligned16(y);
  bn_forward_kernel<<<dim3(splits, c), kBnThreads, 0, (cudaStream_t)stream>>>(
      x, n, c, hw, chunk, vec, partials, gamma, beta, eps, momentum, running_mean, running_var, save_mean, save_invstd, y);
  count_launch();
  return check_launch("bn_forward_kernel");
}

extern "C" int rvb_bn_apply(const float* x, int n, int c, int64_t hw, const float* mean, const float* invstd,
                            const float* gamma, const float* beta, float* y, rvb_stream_t stream) {
  const char* who = "rvb_bn_apply";
  RVB_REQUIRE(x && mean && invstd && y, "%s: null pointer", who);
  const int splits = rvb_bn_splits(n, c, hw);
  int64_t chunk;
  int rc = bn_geometry(who, n, c, hw, splits, &chunk);
  if (rc != RVB_OK) return rc;
  const int vec = (hw % 4 == 0) && aligned16(x) && aligned16(y);
  bn_apply_kernel<<<dim3(splits, c), kBnThreads, 0, (cudaStream_t)stream>>>(x, n, c, hw, chunk, vec, mean, invstd, gamma,
                                                                             beta, y);
  count_launch();
  return check_launch("bn_apply_kernel");
}

extern "C" int rvb_bn_backward(const float* x, const float* dy, int n, int c, int64_t hw, int splits, const double* partials,
                               const float* gamma, const float* mean, const float* invstd, int training, float* dx,
                               float* dgamma, float* dbeta, rvb_stream_t stream) {
  const char* who = "rvb_bn_backward";
  RVB_REQUIRE(x && dy && partials && mean && invstd, "%s: null pointer", who);
  RVB_REQUIRE(dx || dgamma || dbeta, "%s: nothing to compute", who);
  int64_t chunk;
  int rc = bn_geometry(who, n, c, hw, splits, &chunk);
  if (rc != RVB_OK) return rc;
  const int vec = (hw % 4 == 0) && aligned16(x) && aligned16(dy) && aligned16(dx);
  bn_backward_kernel<<<dim3(splits, c), kBnThreads, 0, (cudaStream_t)stream>>>(x, dy, n, c, hw, chunk, vec, partials, gamma,
                                                                                mean, invstd, training, dx, dgamma, dbeta);
  count_launch();
  return check_launch("bn_backward_kernel");
}

// One host call per direction (the caller's network makes ~240 BatchNorm calls per training iteration: the Python /
// ctypes cost per call is what is left of them).  `partials` must hold c * 64 * 2 doubles.
extern "C" int rvb_bn_train_forward(const float* x, int n, int c, int64_t hw, const float* gamma, const float* beta,
                                    float eps, float momentum, float* running_mean, float* running_var, float* save_mean,
                                    float* save_invstd, float* y, double* partials, rvb_stream_t stream) {
  const int splits = rvb_bn_splits(n, c, hw);
  int rc = rvb_bn_reduce(x, nullptr, nullptr, n, c, hw, splits, partials, stream);
  if (rc != RVB_OK) return rc;
  return rvb_bn_forward(x, n, c, hw, splits, partials, gamma, beta, eps, momentum, running_mean, running_var, save_mean,
                        save_invstd, y, stream);
}

extern "C" int rvb_bn_train_backward(const float* x, const float* dy, int n, int c, int64_t hw, const float* gamma,
                                     const float* mean, const float* invstd, int training, float* dx, float* dgamma,
                                     float* dbeta, double* partials, rvb_stream_t stream) {
  const int splits = rvb_bn_splits(n, c, hw);
  int rc = rvb_bn_reduce(x, dy, mean, n, c, hw, splits, partials, stream);
  if (rc != RVB_OK) return rc;
  return rvb_bn_backward(x, dy, n, c, hw, splits, partials, gamma, mean, invstd, training, dx, dgamma, dbeta, stream);
}

// ---- channels_last entry points.  x, y, dy, dx: [n][h][w][c] physical order (torch.channels_last), c % 4 == 0, 16-byte
// aligned; workspace: rvb_bn_nhwc_workspace_bytes(c) bytes, ZEROED once by the caller before its first use.
extern "C" int64_t rvb_bn_nhwc_workspace_bytes(int c) {
  return (int64_t)kBnNhwcMaxBlocks * c * 2 * (int64_t)sizeof(double) + 3 * (int64_t)c * (int64_t)sizeof(float) + 16;
}

extern "C" int rvb_bn_train_forward_nhwc(const float* x, int64_t pixels, int c, const float* gamma, const float* beta,
                                         float eps, float momentum, float* running_mean, float* running_var,
                                         float* save_mean, float* save_invstd, float* y, void* workspace,
                                         rvb_stream_t stream) {
  const char* who = "rvb_bn_train_forward_nhwc";
  RVB_REQUIRE(x && save_mean && save_invstd && y && workspace, "%s: null pointer", who);
  RVB_REQUIRE(pixels > 1, "%s: more than one value per channel is needed in training mode", who);
  RVB_REQUIRE(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta) && aligned16(save_mean) && aligned16(save_invstd),
              "%s: pointers must be 16-byte aligned", who);
  NhwcPlan pl;
  int rc = nhwc_plan(who, pixels, c, &pl);
  if (rc != RVB_OK) return rc;
  const size_t smem = 8 * (size_t)pl.threads * sizeof(double);
  bn_nhwc_reduce_kernel<false><<<pl.blocks, pl.threads, smem, (cudaStream_t)stream>>>(
      x, nullptr, pixels, c, pl.per_block, workspace, eps, momentum, running_mean, running_var, save_mean, save_invstd,
      nullptr, 1, nullptr, nullptr);
  count_launch();
  if ((rc = check_launch("bn_nhwc_reduce_kernel")) != RVB_OK) return rc;
  bn_nhwc_apply_kernel<0><<<pl.blocks, pl.threads, 0, (cudaStream_t)stream>>>(x, nullptr, pixels, c, pl.per_block, save_mean,
                                                                            save_invstd, gamma, beta, workspace, y);
  count_launch();
  return check_launch("bn_nhwc_apply_kernel");
}

extern "C" int rvb_bn_apply_nhwc(const float* x, int64_t pixels, int c, const float* mean, const float* invstd,
                                 const float* gamma, const float* beta, float* y, rvb_stream_t stream) {
  const char* who = "rvb_bn_apply_nhwc";
  RVB_REQUIRE(x && mean && invstd && y, "%s: null pointer", who);
  RVB_REQUIRE(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta) && aligned16(mean) && aligned16(invstd),
              "%s: pointers must be 16-byte aligned", who);
  NhwcPlan pl;
  int rc = nhwc_plan(who, pixels, c, &pl);
  if (rc != RVB_OK) return rc;
  bn_nhwc_apply_kernel<0><<<pl.blocks, pl.threads, 0, (cudaStream_t)stream>>>(x, nullptr, pixels, c, pl.per_block, mean, invstd,
                                                                            gamma, beta, nullptr, y);
  count_launch();
  return check_launch("bn_nhwc_apply_kernel");
}

extern "C" int rvb_bn_train_backward_nhwc(const float* x, const float* dy, int64_t pixels, int c, const float* gamma,
                                          const float* mean, const float* invstd, int training, float* dx, float* dgamma,
                                          float* dbeta, void* workspace, rvb_stream_t stream) {
  const char* who = "rvb_bn_train_backward_nhwc";
  RVB_REQUIRE(x && dy && mean && invstd && workspace, "%s: null pointer", who);
  RVB_REQUIRE(dx || dgamma || dbeta, "%s: nothing to compute", who);
  RVB_REQUIRE(aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(mean) && aligned16(invstd),
              "%s: pointers must be 16-byte aligned", who);
  NhwcPlan pl;
  int rc = nhwc_plan(who, pixels, c, &pl);
  if (rc != RVB_OK) return rc;
  const size_t smem = 8 * (size_t)pl.threads * sizeof(double);
  bn_nhwc_reduce_kernel<true><<<pl.blocks, pl.threads, smem, (cudaStream_t)stream>>>(
      x, dy, pixels, c, pl.per_block, workspace, 0.f, 0.f, nullptr, nullptr, const_cast<float*>(mean),
      const_cast<float*>(invstd), gamma, training, dgamma, dbeta);
  count_launch();
  if ((rc = check_launch("bn_nhwc_reduce_kernel")) != RVB_OK) return rc;
  if (dx == nullptr) return RVB_OK;
  bn_nhwc_apply_kernel<1><<<pl.blocks, pl.threads, 0, (cudaStream_t)stream>>>(x, dy, pixels, c, pl.per_block, mean, invstd, gamma,
                                                                            nullptr, workspace, dx);
  count_launch();
  return check_launch("bn_nhwc_apply_kernel");
}
