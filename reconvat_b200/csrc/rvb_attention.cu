// Local-window multi-head attention of the ReconVAT U-Net (the CALLER of the hot path):
//   model/self_attention_VAT.py:22-91  MutliHeadAttention1D.forward  (same class in UNet_onset.py:22, self_attention.py:6)
//
//   energy[b,l,h,w] = sum_c q[b,l,h,c] * (k[b,l+w-P,h,c] + rel[h,c,w])      rows outside [0,L) are zero (F.pad, bias=False)
//   att = softmax_w(energy);   out[b,l,h,c] = sum_w att[b,l,h,w] * v[b,l+w-P,h,c]
//
// The reference materialises k and v unfolded to (B, L, C, W) -- 73 MB per segment each at C = 916, W = 31 -- and
// autograd keeps them for the backward.  Here a block stages the L_t + W - 1 rows a tile of positions needs in shared
// memory once; nothing of size B*L*C*W ever exists.  The relative-position term does not depend on k: q.rel is a
// small batched GEMM the caller runs on cuBLAS and hands over as `bias` ([B][L][G][W]); likewise dE.rel^T in the
// backward.
//
// Layouts: q, k, v, out, d* : [B][L][G*D] fp32;  att, dE, bias: [B][L][G][W].
// One warp per (position, head).  Shared-memory rows are padded to DS = 4*odd floats (zero filled), so that both
//   energies: lane <-> window slot w, 16-byte loads down row (i+w), and
//   outputs : lane <-> channel quad, 16-byte loads along row (i+w)
// are conflict-free LDS.128 with four FMAs each.
#include "rvb_common.cuh"

namespace rvb {

extern void count_launch();

constexpr int kAttTile = 32;          // positions per block
constexpr int kAttWarps = 32;         // warps per block: one position each
constexpr int kAttMaxQuads = 4;       // channel quads per lane: D <= 512

__host__ __device__ inline int att_row_stride(int D) {
  int q = (D + 3) / 4;
  if ((q & 1) == 0) ++q;              // DS/4 odd: eight consecutive rows hit eight distinct 16-byte bank groups
  return 4 * q;
}

// Stage rows [r0, r0 + n_rows) of head h of a [B][L][C] tensor into smem [n_rows][DS]; rows outside [0, L) and the
// padding columns are zero.
__device__ __forceinline__ void stage_rows(float* __restrict__ dst, const float* __restrict__ src, int b, int h, int r0,
                                           int n_rows, int L, int C, int D, int DS) {
  for (int i = threadIdx.x; i < n_rows * DS; i += blockDim.x) {
    const int r = i / DS, c = i - r * DS;
    const int row = r0 + r;
    dst[i] = (c < D && row >= 0 && row < L) ? __ldg(src + ((int64_t)b * L + row) * C + h * D + c) : 0.f;
  }
}

// sum_c a[c] * b[c] over one padded row pair (a: broadcast row, b: this lane's row)
__device__ __forceinline__ float dot_rows(const float* __restrict__ a, const float* __restrict__ b, int DS) {
  float acc = 0.f;
  for (int c = 0; c < DS; c += 4) {
    const float4 x = *reinterpret_cast<const float4*>(a + c);
    const float4 y = *reinterpret_cast<const float4*>(b + c);
    acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
  }
  return acc;
}

// acc[j] += s * row[4*(lane + 32 j) ..]: lanes <-> channel quads
__device__ __forceinline__ void axpy_row(float4 (&acc)[kAttMaxQuads], float s, const float* __restrict__ row, int lane,
                                         int n_quads) {
#pragma unroll
  for (int j = 0; j < kAttMaxQuads; ++j) {
    const int qd = lane + 32 * j;
    if (qd < n_quads) {
      const float4 x = *reinterpret_cast<const float4*>(row + 4 * qd);
      acc[j].x = fmaf(s, x.x, acc[j].x); acc[j].y = fmaf(s, x.y, acc[j].y);
      acc[j].z = fmaf(s, x.z, acc[j].z); acc[j].w = fmaf(s, x.w, acc[j].w);
    }
  }
}

__device__ __forceinline__ void store_row_quads(float* __restrict__ dst, const float4 (&acc)[kAttMaxQuads], int lane, int D) {
#pragma unroll
  for (int j = 0; j < kAttMaxQuads; ++j) {
    const int c = 4 * (lane + 32 * j);
    if (c < D) dst[c] = acc[j].x;
    if (c + 1 < D) dst[c + 1] = acc[j].y;
    if (c + 2 < D) dst[c + 2] = acc[j].z;
    if (c + 3 < D) dst[c + 3] = acc[j].w;
  }
}

// ---- forward ----------------------------------------------------------------------------
__global__ void __launch_bounds__(kAttWarps * 32)
local_attn_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                      const float* __restrict__ bias, int L, int G, int D, int W, float* __restrict__ out,
                      float* __restrict__ att) {
  extern __shared__ __align__(16) float sm[];
  const int P = (W - 1) / 2, C = G * D, DS = att_row_stride(D);
  const int n_rows = kAttTile + W - 1;
  float* ks = sm;                       // [n_rows][DS]
  float* vs = ks + n_rows * DS;         // [n_rows][DS]
  float* qs = vs + n_rows * DS;         // [kAttTile][DS]
  const int b = blockIdx.z, h = blockIdx.y, l0 = blockIdx.x * kAttTile;
  stage_rows(ks, k, b, h, l0 - P, n_rows, L, C, D, DS);
  stage_rows(vs, v, b, h, l0 - P, n_rows, L, C, D, DS);
  stage_rows(qs, q, b, h, l0, kAttTile, L, C, D, DS);
  __syncthreads();
  const int lane = threadIdx.x & 31, i = threadIdx.x >> 5;
  const int l = l0 + i;
  if (l >= L) return;                                       // warp-uniform, no barrier follows
  const int64_t arow = (((int64_t)b * L + l) * G + h) * W;
  float e = -INFINITY;
  if (lane < W) e = dot_rows(qs + i * DS, ks + (i + lane) * DS, DS) + (bias ? __ldg(bias + arow + lane) : 0.f);
  const float mx = warp_max(e);
  const float pexp = (lane < W) ? expf(e - mx) : 0.f;
  const float a = pexp / warp_sum(pexp);
  if (lane < W) att[arow + lane] = a;
  float4 acc[kAttMaxQuads];
#pragma unroll
  for (int j = 0; j < kAttMaxQuads; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int n_quads = DS >> 2;
  for (int w = 0; w < W; ++w) axpy_row(acc, __shfl_sync(kFull, a, w), vs + (i + w) * DS, lane, n_quads);
  store_row_quads(out + ((int64_t)b * L + l) * C + h * D, acc, lane, D);
}

// ---- backward, part 1: dE (softmax backward) and the k part of dQ --------------------------
//   dAtt[w] = sum_c dOut[c] v[l+w-P][c];  dE[w] = att[w] (dAtt[w] - sum_w' att[w'] dAtt[w']);
//   dQ[c] = sum_w dE[w] k[l+w-P][c]          (+ dE . rel^T, added by the caller)
__global__ void __launch_bounds__(kAttWarps * 32)
local_attn_bwd_q_kernel(const float* __restrict__ dout, const float* __restrict__ att, const float* __restrict__ k,
                        const float* __restrict__ v, int L, int G, int D, int W, float* __restrict__ dE,
                        float* __restrict__ dq) {
  extern __shared__ __align__(16) float sm[];
  const int P = (W - 1) / 2, C = G * D, DS = att_row_stride(D);
  const int n_rows = kAttTile + W - 1;
  float* ks = sm;
  float* vs = ks + n_rows * DS;
  float* gs = vs + n_rows * DS;         // dOut tile [kAttTile][DS]
  const int b = blockIdx.z, h = blockIdx.y, l0 = blockIdx.x * kAttTile;
  stage_rows(ks, k, b, h, l0 - P, n_rows, L, C, D, DS);
  stage_rows(vs, v, b, h, l0 - P, n_rows, L, C, D, DS);
  stage_rows(gs, dout, b, h, l0, kAttTile, L, C, D, DS);
  __syncthreads();
  const int lane = threadIdx.x & 31, i = threadIdx.x >> 5;
  const int l = l0 + i;
  if (l >= L) return;
  const int64_t arow = (((int64_t)b * L + l) * G + h) * W;
  float da = 0.f, a = 0.f;
  if (lane < W) {
    a = __ldg(att + arow + lane);
    da = dot_rows(gs + i * DS, vs + (i + lane) * DS, DS);
  }
  const float s = warp_sum(a * da);
  const float de = a * (da - s);
  if (lane < W) dE[arow + lane] = de;
  float4 acc[kAttMaxQuads];
#pragma unroll
  for (int j = 0; j < kAttMaxQuads; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int n_quads = DS >> 2;
  for (int w = 0; w < W; ++w) axpy_row(acc, __shfl_sync(kFull, de, w), ks + (i + w) * DS, lane, n_quads);
  store_row_quads(dq + ((int64_t)b * L + l) * C + h * D, acc, lane, D);
}

// ---- backward, part 2: dK and dV in gather form -------------------------------------------
//   source row m is slot w of position l = m - w + P:
//   dK[m][c] = sum_w dE[l][w] q[l][c],   dV[m][c] = sum_w att[l][w] dOut[l][c]      (l in [0, L))
__global__ void __launch_bounds__(kAttWarps * 32)
local_attn_bwd_kv_kernel(const float* __restrict__ q, const float* __restrict__ dout, const float* __restrict__ att,
                         const float* __restrict__ dE, int L, int G, int D, int W, float* __restrict__ dk,
                         float* __restrict__ dv) {
  extern __shared__ __align__(16) float sm[];
  const int P = (W - 1) / 2, C = G * D, DS = att_row_stride(D);
  const int n_rows = kAttTile + W - 1;
  float* qs = sm;                       // rows l in [m0 - P, m0 + kAttTile + P)
  float* gs = qs + n_rows * DS;
  const int b = blockIdx.z, h = blockIdx.y, m0 = blockIdx.x * kAttTile;
  stage_rows(qs, q, b, h, m0 - P, n_rows, L, C, D, DS);
  stage_rows(gs, dout, b, h, m0 - P, n_rows, L, C, D, DS);
  __syncthreads();
  const int lane = threadIdx.x & 31, i = threadIdx.x >> 5;
  const int m = m0 + i;
  if (m >= L) return;
  // lane w holds the two scalars of (l = m - w + P, slot w)
  float de = 0.f, a = 0.f;
  const int lw = m - lane + P;
  if (lane < W && lw >= 0 && lw < L) {
    const int64_t idx = (((int64_t)b * L + lw) * G + h) * W + lane;
    de = __ldg(dE + idx);
    a = __ldg(att + idx);
  }
  float4 ak[kAttMaxQuads], av[kAttMaxQuads];
#pragma unroll
  for (int j = 0; j < kAttMaxQuads; ++j) { ak[j] = make_float4(0.f, 0.f, 0.f, 0.f); av[j] = ak[j]; }
  const int n_quads = DS >> 2;
  for (int w = 0; w < W; ++w) {
    const int r = i - w + 2 * P;                            // smem row of l = m - w + P  (row 0 <-> l = m0 - P)
    axpy_row(ak, __shfl_sync(kFull, de, w), qs + r * DS, lane, n_quads);
    axpy_row(av, __shfl_sync(kFull, a, w), gs + r * DS, lane, n_quads);
  }
  store_row_quads(dk + ((int64_t)b * L + m) * C + h * D, ak, lane, D);
  store_row_quads(dv + ((int64_t)b * L + m) * C + h * D, av, lane, D);
}

static int attn_check(const char* who, int B, int L, int G, int D, int W, int n_row_sets, int n_tile_sets,
                      size_t* smem_bytes) {
  RVB_REQUIRE(B > 0 && L > 0 && G > 0 && D > 0, "%s: bad shape", who);
  RVB_REQUIRE(B <= 65535 && G <= 65535, "%s: batch / heads too large for the grid", who);
  RVB_REQUIRE(W >= 1 && W <= 32 && (W & 1), "%s: window %d must be odd and <= 32", who, W);
  RVB_REQUIRE(att_row_stride(D) <= 4 * 32 * kAttMaxQuads, "%s: head dimension %d > %d", who, D, 4 * 32 * kAttMaxQuads - 4);
  *smem_bytes = (size_t)(n_row_sets * (kAttTile + W - 1) + n_tile_sets * kAttTile) * att_row_stride(D) * sizeof(float);
  RVB_REQUIRE(*smem_bytes <= 220 * 1024, "%s: head dimension %d with window %d needs %zu bytes of shared memory", who, D, W,
              *smem_bytes);
  return RVB_OK;
}

}  // namespace rvb

using namespace rvb;

extern "C" int rvb_local_attn_fwd(const float* q, const float* k, const float* v, const float* bias, int B, int L, int G,
                                  int D, int W, float* out, float* att, rvb_stream_t stream) {
  RVB_REQUIRE(q && k && v && out && att, "rvb_local_attn_fwd: null pointer");
  size_t smem;
  int rc = attn_check("rvb_local_attn_fwd", B, L, G, D, W, 2, 1, &smem);
  if (rc != RVB_OK) return rc;
  RVB_CUDA(cudaFuncSetAttribute(local_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((L + kAttTile - 1) / kAttTile), (unsigned)G, (unsigned)B);
  local_attn_fwd_kernel<<<grid, kAttWarps * 32, smem, (cudaStream_t)stream>>>(q, k, v, bias, L, G, D, W, out, att);
  count_launch();
  return check_launch("local_attn_fwd_kernel");
}

extern "C" int rvb_local_attn_bwd_q(const float* dout, const float* att, const float* k, const float* v, int B, int L,
                                    int G, int D, int W, float* dE, float* dq, rvb_stream_t stream) {
  RVB_REQUIRE(dout && att && k && v && dE && dq, "rvb_local_attn_bwd_q: null pointer");
  size_t smem;
  int rc = attn_check("rvb_local_attn_bwd_q", B, L, G, D, W, 2, 1, &smem);
  if (rc != RVB_OK) return rc;
  RVB_CUDA(cudaFuncSetAttribute(local_attn_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((L + kAttTile - 1) / kAttTile), (unsigned)G, (unsigned)B);
  local_attn_bwd_q_kernel<<<grid, kAttWarps * 32, smem, (cudaStream_t)stream>>>(dout, att, k, v, L, G, D, W, dE, dq);
  count_launch();
  return check_launch("local_attn_bwd_q_kernel");
}

extern "C" int rvb_local_attn_bwd_kv(const float* q, const float* dout, const float* att, const float* dE, int B, int L,
                                     int G, int D, int W, float* dk, float* dv, rvb_stream_t stream) {
  RVB_REQUIRE(q && dout && att && dE && dk && dv, "rvb_local_attn_bwd_kv: null pointer");
  size_t smem;
  int rc = attn_check("rvb_local_attn_bwd_kv", B, L, G, D, W, 2, 0, &smem);
  if (rc != RVB_OK) return rc;
  RVB_CUDA(cudaFuncSetAttribute(local_attn_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((L + kAttTile - 1) / kAttTile), (unsigned)G, (unsigned)B);
  local_attn_bwd_kv_kernel<<<grid, kAttWarps * 32, smem, (cudaStream_t)stream>>>(q, dout, att, dE, L, G, D, W, dk, dv);
  count_launch();
  return check_launch("local_attn_bwd_kv_kernel");
}


// ---------------------------------------------------------------- operand planes of rvb_gemm_nt_tf32x3
// x [rows][cols] (row stride ld)  ->  tf32 hi / lo planes, hi = tf32(x), lo = tf32(x - hi):
//   transpose == 0:  planes[r][col0 + c] = split(x[r][c])      plane row stride out_ld; columns [col0 + cols, out_ld) of the
//                    LAST block written by a call are the caller's to zero (rvb_split_tf32 zeroes [col0 + cols, zero_to))
//   transpose == 1:  planes[c][row0 + r] = split(x[r][c])      (the contraction runs over x's ROWS: dW = dY^T X)
namespace rvb {

__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ x, int64_t rows, int cols, int64_t ld, float* __restrict__ hi,
                  float* __restrict__ lo, int64_t out_ld, int col0, int zero_to, int vec) {
  // vec: 16-byte loads and stores (every row start, col0 and the widths are multiples of four floats)
  if (vec) {
    const int c4 = blockIdx.x * blockDim.x + threadIdx.x;           // quad index inside [col0, zero_to)
    if (c4 >= (zero_to - col0) >> 2) return;
    const int c = c4 << 2;
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c + 4 <= cols) v = __ldg(reinterpret_cast<const float4*>(x + r * ld + c));
      else {
        if (c < cols) v.x = __ldg(x + r * ld + c);
        if (c + 1 < cols) v.y = __ldg(x + r * ld + c + 1);
        if (c + 2 < cols) v.z = __ldg(x + r * ld + c + 2);
      }
      const float4 h = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
      *reinterpret_cast<float4*>(hi + r * out_ld + col0 + c) = h;
      *reinterpret_cast<float4*>(lo + r * out_ld + col0 + c) =
          make_float4(to_tf32(v.x - h.x), to_tf32(v.y - h.y), to_tf32(v.z - h.z), to_tf32(v.w - h.w));
    }
    return;
  }
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= zero_to - col0) return;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    const float v = (c < cols) ? __ldg(x + r * ld + c) : 0.f;
    const float h = to_tf32(v);
    hi[r * out_ld + col0 + c] = h;
    lo[r * out_ld + col0 + c] = to_tf32(v - h);
  }
}

// 32 x 32 tiles through shared memory: coalesced reads along x's columns, coalesced writes along x's rows
__global__ void __launch_bounds__(256)
split_tf32_transpose_kernel(const float* __restrict__ x, int64_t rows, int cols, int64_t ld, float* __restrict__ hi,
                            float* __restrict__ lo, int64_t out_ld, int64_t row0, int64_t zero_to) {
  __shared__ float tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.y * 32;
  const int c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int64_t r = r0 + ty + j;
    const int c = c0 + tx;
    tile[ty + j][tx] = (r < rows && c < cols) ? __ldg(x + r * ld + c) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int c = c0 + ty + j;                                      // output row
    const int64_t r = r0 + tx;                                      // output column (before row0)
    if (c < cols && row0 + r < zero_to) {
      const float v = tile[tx][ty + j];                             // zero past `rows`
      const float h = to_tf32(v);
      hi[(int64_t)c * out_ld + row0 + r] = h;
      lo[(int64_t)c * out_ld + row0 + r] = to_tf32(v - h);
    }
  }
}

}  // namespace rvb

extern "C" int rvb_split_tf32(const float* x, int64_t rows, int cols, int64_t ld, int transpose, float* hi, float* lo,
                              int64_t out_ld, int64_t offset, int64_t zero_to, rvb_stream_t stream) {
  RVB_REQUIRE(x && hi && lo, "rvb_split_tf32: null pointer");
  RVB_REQUIRE(rows > 0 && cols > 0 && ld >= cols && offset >= 0, "rvb_split_tf32: bad shape");
  const int64_t extent = transpose ? rows : cols;                   // what runs along the plane's columns
  RVB_REQUIRE(zero_to >= offset + extent && zero_to <= out_ld, "rvb_split_tf32: need offset + extent <= zero_to <= out_ld");
  if (!transpose) {
    RVB_REQUIRE(zero_to - offset < (1ll << 31), "rvb_split_tf32: too wide");
    const auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    const int vec = a16(x) && a16(hi) && a16(lo) && (ld & 3) == 0 && (out_ld & 3) == 0 && (offset & 3) == 0 &&
                    ((zero_to - offset) & 3) == 0;
    const int64_t width = vec ? (zero_to - offset) >> 2 : zero_to - offset;
    const unsigned bx = (unsigned)((width + 255) / 256);
    // few, fat blocks along the rows: every thread streams a column quad down its share of the rows
    const int64_t by = rows < 2048 ? rows : 2048;
    dim3 grid(bx, (unsigned)by);
    rvb::split_tf32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, ld, hi, lo, out_ld, (int)offset, (int)zero_to, vec);
  } else {
    const int64_t r_tiles = (zero_to - offset + 31) / 32;
    RVB_REQUIRE(r_tiles <= 65535, "rvb_split_tf32: too many rows for one launch");
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)r_tiles);
    rvb::split_tf32_transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, ld, hi, lo, out_ld, offset, zero_to);
  }
  rvb::count_launch();
  return rvb::check_launch("split_tf32_kernel");
}
