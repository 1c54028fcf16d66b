// Local-window multi-head attention of the ReconVAT U-Net (the CALLER of the hot path):
//   model/self_attention_VAT.py:22-91  MutliHeadAttention1D.forward  (same class in UNet_onset.py:22, self_attention.py:6)
//
//   energy[b,l,h,w] = sum_c q[b,l,h,c] * (k[b,l+w-P,h,c] + rel[h,c,w])      rows outside [0,L) are zero (F.pad, bias=False)
//   att = softmax_w(energy);   out[b,l,h,c] = sum_w att[b,l,h,w] * v[b,l+w-P,h,c]
//
// The reference materialises k and v unfolded to (B, L, C, W) -- 73 MB per segment each at C = 916, W = 31 -- and
// autograd keeps them for the backward.  Here a block stages the L_t + W - 1 rows a tile of positions needs in shared
// memory once; nothing of size B*L*C*W ever exists.
//
// Layouts: q, k, v, out, d* : [B][L][G*D] fp32;  rel: [G*D][W] and its per-head transpose relT: [G][W][D];
// att, dE: [B][L][G][W].
// One warp per (position, head).  Energies: lane <-> window slot w (row stride D is odd in practice, so the 32 lanes
// hit 32 banks), loop over the D channels.  Outputs: lane <-> channel (c = lane, lane+32, ...), loop over the window.
#include "rvb_common.cuh"

namespace rvb {

extern void count_launch();

constexpr int kAttTile = 16;          // positions per block
constexpr int kAttWarps = 16;         // warps per block: one position each
constexpr int kAttMaxDPerLane = 16;   // D <= 512

// Stage rows [r0, r0 + n_rows) of head h of a [B][L][C] tensor into smem [n_rows][D]; rows outside [0, L) are zero.
__device__ __forceinline__ void stage_rows(float* __restrict__ dst, const float* __restrict__ src, int b, int h, int r0,
                                           int n_rows, int L, int C, int D) {
  for (int i = threadIdx.x; i < n_rows * D; i += blockDim.x) {
    const int r = i / D, c = i - r * D;
    const int row = r0 + r;
    dst[i] = (row >= 0 && row < L) ? __ldg(src + ((int64_t)b * L + row) * C + h * D + c) : 0.f;
  }
}

// ---- forward ----------------------------------------------------------------------------
__global__ void __launch_bounds__(kAttWarps * 32)
local_attn_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                      const float* __restrict__ rel, int L, int G, int D, int W, float* __restrict__ out,
                      float* __restrict__ att) {
  extern __shared__ __align__(16) float sm[];
  const int P = (W - 1) / 2, C = G * D;
  const int n_rows = kAttTile + W - 1;
  float* ks = sm;                       // [n_rows][D]
  float* vs = ks + n_rows * D;          // [n_rows][D]
  float* qs = vs + n_rows * D;          // [kAttTile][D]
  const int b = blockIdx.z, h = blockIdx.y, l0 = blockIdx.x * kAttTile;
  stage_rows(ks, k, b, h, l0 - P, n_rows, L, C, D);
  stage_rows(vs, v, b, h, l0 - P, n_rows, L, C, D);
  stage_rows(qs, q, b, h, l0, kAttTile, L, C, D);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* relh = rel + (int64_t)h * D * W;
  for (int i = warp; i < kAttTile; i += kAttWarps) {
    const int l = l0 + i;
    if (l >= L) break;                                      // warp-uniform
    const float* qi = qs + i * D;
    // lane w: energy of window slot w
    float e = 0.f;
    if (lane < W) {
      const float* kw = ks + (i + lane) * D;
      for (int c = 0; c < D; ++c) e = fmaf(qi[c], kw[c] + __ldg(relh + c * W + lane), e);
    } else {
      e = -INFINITY;
    }
    const float mx = warp_max(e);
    float pexp = (lane < W) ? expf(e - mx) : 0.f;
    const float den = warp_sum(pexp);
    const float a = pexp / den;
    if (lane < W) att[(((int64_t)b * L + l) * G + h) * W + lane] = a;
    // lane <-> channel
    float acc[kAttMaxDPerLane];
#pragma unroll
    for (int j = 0; j < kAttMaxDPerLane; ++j) acc[j] = 0.f;
    for (int w = 0; w < W; ++w) {
      const float aw = __shfl_sync(kFull, a, w);
      const float* vw = vs + (i + w) * D;
#pragma unroll
      for (int j = 0; j < kAttMaxDPerLane; ++j) {
        const int c = lane + 32 * j;
        if (c < D) acc[j] = fmaf(aw, vw[c], acc[j]);
      }
    }
    float* o = out + ((int64_t)b * L + l) * C + h * D;
#pragma unroll
    for (int j = 0; j < kAttMaxDPerLane; ++j) {
      const int c = lane + 32 * j;
      if (c < D) o[c] = acc[j];
    }
  }
}

// ---- backward, part 1: dE (softmax backward) and dQ --------------------------------------
//   dAtt[w] = sum_c dOut[c] v[l+w-P][c];  dE[w] = att[w] (dAtt[w] - sum_w' att[w'] dAtt[w']);
//   dQ[c] = sum_w dE[w] (k[l+w-P][c] + rel[c][w])
__global__ void __launch_bounds__(kAttWarps * 32)
local_attn_bwd_q_kernel(const float* __restrict__ dout, const float* __restrict__ att, const float* __restrict__ k,
                        const float* __restrict__ v, const float* __restrict__ relT, int L, int G, int D, int W,
                        float* __restrict__ dE, float* __restrict__ dq) {
  extern __shared__ __align__(16) float sm[];
  const int P = (W - 1) / 2, C = G * D;
  const int n_rows = kAttTile + W - 1;
  float* ks = sm;
  float* vs = ks + n_rows * D;
  float* gs = vs + n_rows * D;          // dOut tile [kAttTile][D]
  const int b = blockIdx.z, h = blockIdx.y, l0 = blockIdx.x * kAttTile;
  stage_rows(ks, k, b, h, l0 - P, n_rows, L, C, D);
  stage_rows(vs, v, b, h, l0 - P, n_rows, L, C, D);
  stage_rows(gs, dout, b, h, l0, kAttTile, L, C, D);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* relh = relT + (int64_t)h * W * D;           // [W][D]: lanes (channels) read consecutive floats
  for (int i = warp; i < kAttTile; i += kAttWarps) {
    const int l = l0 + i;
    if (l >= L) break;
    const float* gi = gs + i * D;
    const int64_t arow = (((int64_t)b * L + l) * G + h) * W;
    float da = 0.f, a = 0.f;
    if (lane < W) {
      a = __ldg(att + arow + lane);
      const float* vw = vs + (i + lane) * D;
      for (int c = 0; c < D; ++c) da = fmaf(gi[c], vw[c], da);
    }
    const float s = warp_sum(a * da);
    const float de = a * (da - s);
    if (lane < W) dE[arow + lane] = de;
    float acc[kAttMaxDPerLane];
#pragma unroll
    for (int j = 0; j < kAttMaxDPerLane; ++j) acc[j] = 0.f;
    for (int w = 0; w < W; ++w) {
      const float dw = __shfl_sync(kFull, de, w);
      const float* kw = ks + (i + w) * D;
#pragma unroll
      for (int j = 0; j < kAttMaxDPerLane; ++j) {
        const int c = lane + 32 * j;
        if (c < D) acc[j] = fmaf(dw, kw[c] + __ldg(relh + w * D + c), acc[j]);
      }
    }
    float* o = dq + ((int64_t)b * L + l) * C + h * D;
#pragma unroll
    for (int j = 0; j < kAttMaxDPerLane; ++j) {
      const int c = lane + 32 * j;
      if (c < D) o[c] = acc[j];
    }
  }
}

// ---- backward, part 2: dK and dV in gather form -------------------------------------------
//   source row m is slot w of position l = m - w + P:
//   dK[m][c] = sum_w dE[l][w] q[l][c],   dV[m][c] = sum_w att[l][w] dOut[l][c]      (l in [0, L))
__global__ void __launch_bounds__(kAttWarps * 32)
local_attn_bwd_kv_kernel(const float* __restrict__ q, const float* __restrict__ dout, const float* __restrict__ att,
                         const float* __restrict__ dE, int L, int G, int D, int W, float* __restrict__ dk,
                         float* __restrict__ dv) {
  extern __shared__ __align__(16) float sm[];
  const int P = (W - 1) / 2, C = G * D;
  const int n_rows = kAttTile + W - 1;
  float* qs = sm;                       // rows l in [m0 - P, m0 + kAttTile + P)
  float* gs = qs + n_rows * D;
  const int b = blockIdx.z, h = blockIdx.y, m0 = blockIdx.x * kAttTile;
  stage_rows(qs, q, b, h, m0 - P, n_rows, L, C, D);
  stage_rows(gs, dout, b, h, m0 - P, n_rows, L, C, D);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = warp; i < kAttTile; i += kAttWarps) {
    const int m = m0 + i;
    if (m >= L) break;
    // lane w holds the two scalars of (l = m - w + P, slot w)
    float de = 0.f, a = 0.f;
    const int lw = m - lane + P;
    if (lane < W && lw >= 0 && lw < L) {
      const int64_t idx = (((int64_t)b * L + lw) * G + h) * W + lane;
      de = __ldg(dE + idx);
      a = __ldg(att + idx);
    }
    float ak[kAttMaxDPerLane], av[kAttMaxDPerLane];
#pragma unroll
    for (int j = 0; j < kAttMaxDPerLane; ++j) { ak[j] = 0.f; av[j] = 0.f; }
    for (int w = 0; w < W; ++w) {
      const float dw = __shfl_sync(kFull, de, w), aw = __shfl_sync(kFull, a, w);
      const int r = i - w + 2 * P;                          // smem row of l = m - w + P  (row 0 <-> l = m0 - P)
      const float* qr = qs + r * D;
      const float* gr = gs + r * D;
#pragma unroll
      for (int j = 0; j < kAttMaxDPerLane; ++j) {
        const int c = lane + 32 * j;
        if (c < D) {
          ak[j] = fmaf(dw, qr[c], ak[j]);
          av[j] = fmaf(aw, gr[c], av[j]);
        }
      }
    }
    float* ok = dk + ((int64_t)b * L + m) * C + h * D;
    float* ov = dv + ((int64_t)b * L + m) * C + h * D;
#pragma unroll
    for (int j = 0; j < kAttMaxDPerLane; ++j) {
      const int c = lane + 32 * j;
      if (c < D) { ok[c] = ak[j]; ov[c] = av[j]; }
    }
  }
}

static int attn_check(const char* who, int B, int L, int G, int D, int W, size_t smem_floats, size_t* smem_bytes) {
  RVB_REQUIRE(B > 0 && L > 0 && G > 0 && D > 0, "%s: bad shape", who);
  RVB_REQUIRE(B <= 65535 && G <= 65535, "%s: batch / heads too large for the grid", who);
  RVB_REQUIRE(W >= 1 && W <= 32 && (W & 1), "%s: window %d must be odd and <= 32", who, W);
  RVB_REQUIRE(D <= 32 * kAttMaxDPerLane, "%s: head dimension %d > %d", who, D, 32 * kAttMaxDPerLane);
  *smem_bytes = smem_floats * sizeof(float);
  RVB_REQUIRE(*smem_bytes <= 200 * 1024, "%s: head dimension %d with window %d needs %zu bytes of shared memory", who, D, W,
              *smem_bytes);
  return RVB_OK;
}

}  // namespace rvb

using namespace rvb;

extern "C" int rvb_local_attn_fwd(const float* q, const float* k, const float* v, const float* rel, int B, int L, int G,
                                  int D, int W, float* out, float* att, rvb_stream_t stream) {
  RVB_REQUIRE(q && k && v && rel && out && att, "rvb_local_attn_fwd: null pointer");
  size_t smem;
  int rc = attn_check("rvb_local_attn_fwd", B, L, G, D, W, (size_t)(2 * (kAttTile + W - 1) + kAttTile) * D, &smem);
  if (rc != RVB_OK) return rc;
  if (smem > 48 * 1024)
    RVB_CUDA(cudaFuncSetAttribute(local_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((L + kAttTile - 1) / kAttTile), (unsigned)G, (unsigned)B);
  local_attn_fwd_kernel<<<grid, kAttWarps * 32, smem, (cudaStream_t)stream>>>(q, k, v, rel, L, G, D, W, out, att);
  count_launch();
  return check_launch("local_attn_fwd_kernel");
}

extern "C" int rvb_local_attn_bwd_q(const float* dout, const float* att, const float* k, const float* v,
                                    const float* relT, int B, int L, int G, int D, int W, float* dE, float* dq,
                                    rvb_stream_t stream) {
  RVB_REQUIRE(dout && att && k && v && relT && dE && dq, "rvb_local_attn_bwd_q: null pointer");
  size_t smem;
  int rc = attn_check("rvb_local_attn_bwd_q", B, L, G, D, W, (size_t)(2 * (kAttTile + W - 1) + kAttTile) * D, &smem);
  if (rc != RVB_OK) return rc;
  if (smem > 48 * 1024)
    RVB_CUDA(cudaFuncSetAttribute(local_attn_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((L + kAttTile - 1) / kAttTile), (unsigned)G, (unsigned)B);
  local_attn_bwd_q_kernel<<<grid, kAttWarps * 32, smem, (cudaStream_t)stream>>>(dout, att, k, v, relT, L, G, D, W, dE, dq);
  count_launch();
  return check_launch("local_attn_bwd_q_kernel");
}

extern "C" int rvb_local_attn_bwd_kv(const float* q, const float* dout, const float* att, const float* dE, int B, int L,
                                     int G, int D, int W, float* dk, float* dv, rvb_stream_t stream) {
  RVB_REQUIRE(q && dout && att && dE && dk && dv, "rvb_local_attn_bwd_kv: null pointer");
  size_t smem;
  int rc = attn_check("rvb_local_attn_bwd_kv", B, L, G, D, W, (size_t)2 * (kAttTile + W - 1) * D, &smem);
  if (rc != RVB_OK) return rc;
  if (smem > 48 * 1024)
    RVB_CUDA(cudaFuncSetAttribute(local_attn_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((L + kAttTile - 1) / kAttTile), (unsigned)G, (unsigned)B);
  local_attn_bwd_kv_kernel<<<grid, kAttWarps * 32, smem, (cudaStream_t)stream>>>(q, dout, att, dE, L, G, D, W, dk, dv);
  count_launch();
  return check_launch("local_attn_bwd_kv_kernel");
}
