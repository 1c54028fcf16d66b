// K1: the STFT as one dense contraction on 5th-generation tensor cores (sm_100a).
//
//   D[frame, col] = sum_n  p[hop*frame + n] * basis[col, n]          (model/Spectrogram.py:219-220)
//
// * M = frames (B*T), N = basis rows (cos|sin interleaved per 128-bin tile), K = n_fft.
// * 3xTF32: operands are pre-split into tf32 hi/lo planes; each K-step issues hi*hi + hi*lo + lo*hi
//   with fp32 accumulation in TMEM (error ~2^-21 per product, inside the 1e-4 log-Mel budget; a
//   single TF32 pass is 2.3e-3 off on white noise and 1.7e-2 on tonal input -- see DESIGN.md).
// * A is never materialised as frames: hop | n_fft-block arithmetic makes K-slice [kk, kk+32) of frame
//   t the 32 floats at column kk%hop of row t + kk/hop of the hop-blocked signal plane, so a
//   128-frame x 32-sample operand tile is ONE 2-D TMA box over non-overlapping rows.
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma issuer,
//   warps 2..5 = epilogue (TMEM -> registers -> re^2+im^2 etc. -> coalesced global stores).
//   smem ring (full/empty mbarriers) between producer and MMA; two 256-column TMEM accumulators
//   (tmem_full/tmem_empty mbarriers) so the epilogue of unit i overlaps the MMAs of unit i+1.
// * Persistent: grid = min(#units, #SMs); unit = (128-frame tile, 256-column tile), n fastest.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdlib>
#include <cstring>
#include <map>
#include <type_traits>
#include <mutex>
#include <tuple>

#include "rvb_common.cuh"

namespace rvb {

extern void count_launch();

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 32;                       // fp32 elements = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 8;                         // tf32
constexpr int STAGES = 2;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 4;   // 16 KB
constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 4;   // 32 KB
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;   // hi+lo of both operands: 96 KB
constexpr int BAR_BYTES = 128;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;   // + slack for 1024-byte alignment
constexpr int ACC_COLS = BLOCK_N;                 // fp32 accumulator columns per stage
constexpr int TMEM_COLS = 512;
constexpr int NUM_THREADS = 192;
constexpr int EPI_WARP0 = 2;

constexpr int kEpiRawGemm = 0x100;                // internal epilogue code: store the accumulator (plain GEMM)

struct GemmParams {
  int64_t ldc;                                    // kEpiRawGemm: row stride of C
  int64_t split_stride;                           // kEpiRawGemm: elements between the partial results of a split contraction
  int k_split;                                    // units per (m, n) tile along the contraction (1 for the STFT)
  int n_seg, rows_per_seg, hop, n_frames, n_fft;
  int tiles_per_seg, m_tiles, n_tiles;
  int epilogue, n_out_bins, n_store_bins;
  float power;
  float* out0;
  int* dbg_status;
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug must become a trap, never a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* dbg_status, int code) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0 = 0;
  for (uint32_t it = 0;; ++it) {
    if (mbar_try_wait(bar, parity)) return;
    if ((it & 0x3ff) == 0x3ff) {
      const uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000ull) {     // 4 s
        if (dbg_status) atomicExch(dbg_status, code);
        __threadfence_system();
        __trap();
      }
    }
  }
}
// The same with cluster-scope acquire: the arrivals come from threads of the peer CTA that wrote ITS shared memory.
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int* dbg_status, int code) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  uint64_t t0 = 0;
  for (uint32_t it = 0;; ++it) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    if ((it & 0x3ff) == 0x3ff) {
      const uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000ull) {     // 4 s
        if (dbg_status) atomicExch(dbg_status, code);
        __threadfence_system();
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint32_t smem_dst, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (tile base 1024-byte aligned):
//   bits [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major, 1) | [32,46) SBO>>4 = 1024>>4
//   | [46,48) version = 1 (sm_100) | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  const uint32_t lo = ((smem_addr >> 4) & 0x3fffu) | (1u << 16);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return (uint64_t)lo | ((uint64_t)hi << 32);
}
// Instruction descriptor: D=f32 (bits 4-5 = 1), A and B formats at bits 7-9 / 10-12 (0 = f16, 2 = tf32), both
// K-major, N>>3 at bits 17-22, M>>4 at bits 24-28.
constexpr uint32_t FMT_F16 = 0, FMT_TF32 = 2;
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, uint32_t fmt) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) { return make_idesc(m, n, FMT_TF32); }
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- kernel
__global__ void __launch_bounds__(NUM_THREADS, 1)
stft_gemm_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                 const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                 const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto s_a_hi = [&](int s) { return smem_base + s * STAGE_BYTES; };
  auto s_a_lo = [&](int s) { return smem_base + s * STAGE_BYTES + A_TILE_BYTES; };
  auto s_b_hi = [&](int s) { return smem_base + s * STAGE_BYTES + 2 * A_TILE_BYTES; };
  auto s_b_lo = [&](int s) { return smem_base + s * STAGE_BYTES + 2 * A_TILE_BYTES + B_TILE_BYTES; };
  auto bar_full = [&](int s) { return bar_base + 8 * s; };
  auto bar_empty = [&](int s) { return bar_base + 8 * (STAGES + s); };
  auto bar_tmem_full = [&](int a) { return bar_base + 8 * (2 * STAGES + a); };
  auto bar_tmem_empty = [&](int a) { return bar_base + 8 * (2 * STAGES + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8 * (2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;   // warp-uniform
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_a_lo);
    tma_prefetch_desc(&tm_b_hi);
    tma_prefetch_desc(&tm_b_lo);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tmem_full(a), 1);
      mbar_init(bar_tmem_empty(a), 4);          // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  __syncthreads();                                     // mbarrier words initialised before the allocator (another warp) writes beside them
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr) : "memory");

  // unit = (m tile, n tile, slice of the contraction); k_split == 1 everywhere but in the split-K plain GEMM
  const int n_units = p.m_tiles * p.n_tiles * p.k_split;
  const int num_kb_all = p.n_fft / BLOCK_K;
  const int kb_per = (num_kb_all + p.k_split - 1) / p.k_split;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int ks = unit % p.k_split, mn = unit / p.k_split;
        const int m_tile = mn / p.n_tiles, n_tile = mn - m_tile * p.n_tiles;
        const int b = m_tile / p.tiles_per_seg;
        const int t0 = (m_tile - b * p.tiles_per_seg) * BLOCK_M;
        const int row0 = b * p.rows_per_seg + t0;
        const int kb_end = min(num_kb_all, (ks + 1) * kb_per);
        for (int kb = ks * kb_per; kb < kb_end; ++kb) {
          mbar_wait(bar_empty(stage), phase ^ 1u, p.dbg_status, 1);
          mbar_expect_tx(bar_full(stage), STAGE_BYTES);
          const int kk = kb * BLOCK_K;
          const int col = kk % p.hop, row = row0 + kk / p.hop;
          tma_load_2d(&tm_a_hi, s_a_hi(stage), bar_full(stage), col, row);
          tma_load_2d(&tm_a_lo, s_a_lo(stage), bar_full(stage), col, row);
          tma_load_2d(&tm_b_hi, s_b_hi(stage), bar_full(stage), kk, n_tile * BLOCK_N);
          tma_load_2d(&tm_b_lo, s_b_lo(stage), bar_full(stage), kk, n_tile * BLOCK_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BLOCK_M, BLOCK_N);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        mbar_wait(bar_tmem_empty(acc), acc_phase ^ 1u, p.dbg_status, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
        const int kb0 = (unit % p.k_split) * kb_per;
        const int num_kb = min(num_kb_all, kb0 + kb_per) - kb0;      // >= 1: the host never makes empty slices
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_full(stage), phase, p.dbg_status, 3);
          tc_fence_after();
          const uint64_t da_hi = make_sw128_desc(s_a_hi(stage));
          const uint64_t da_lo = make_sw128_desc(s_a_lo(stage));
          const uint64_t db_hi = make_sw128_desc(s_b_hi(stage));
          const uint64_t db_lo = make_sw128_desc(s_b_lo(stage));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)(k * UMMA_K * 4 >> 4);      // +32 bytes inside the swizzle row
            umma_tf32(d_tmem, da_hi + adv, db_hi + adv, idesc, (kb | k) != 0);
            umma_tf32(d_tmem, da_hi + adv, db_lo + adv, idesc, 1u);
            umma_tf32(d_tmem, da_lo + adv, db_hi + adv, idesc, 1u);
          }
          umma_commit(bar_empty(stage));            // smem slot reusable once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(bar_tmem_full(acc));            // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps =====================
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      const int ks = unit % p.k_split, mn = unit / p.k_split;
      const int m_tile = mn / p.n_tiles, n_tile = mn - m_tile * p.n_tiles;
      const int b = m_tile / p.tiles_per_seg;
      const int t = (m_tile - b * p.tiles_per_seg) * BLOCK_M + row;
      const bool t_ok = t < p.n_frames;
      mbar_wait(bar_tmem_full(acc), acc_phase, p.dbg_status, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_COLS);
      if (p.epilogue == kEpiRawGemm) {
        // plain GEMM (rvb_gemm_nt_tf32x3): C[t][n_tile * 256 + j] = accumulator column j; n_out_bins = N, power = ldc
        float* crow = p.out0 + (int64_t)ks * p.split_stride + (int64_t)t * p.ldc + n_tile * BLOCK_N;
        const bool vec = (p.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out0) & 15u) == 0;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
          const int n0 = n_tile * BLOCK_N + c * 32;
          if (t_ok && n0 < p.n_out_bins) {
            if (vec && n0 + 32 <= p.n_out_bins) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                reinterpret_cast<float4*>(crow + c * 32)[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                                          __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (n0 + i < p.n_out_bins) crow[c * 32 + i] = __uint_as_float(v[i]);
            }
          }
        }
      } else {
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t re[32], im[32];
        tmem_ld32(taddr + c * 32, re);
        tmem_ld32(taddr + 128 + c * 32, im);
        tmem_ld_wait();
        const int k0 = n_tile * 128 + c * 32;
        if (t_ok)
          stft_store_chunk(p.epilogue, p.power, re, im, 1.f, 0.f, p.out0, b, k0, t, p.n_out_bins, p.n_store_bins, p.n_frames);
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tmem_empty(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------- folded kernel (K1f)
// Same pipeline skeleton; differences from stft_gemm_kernel:
//  * the K loop runs two chains back to back: chain 0 = e x folded-cos -> accumulator columns [0,128) (re),
//    chain 1 = o x folded-sin -> columns [128,256) (im); each chain is K = n_fft/2 long, so the MMA work is
//    half of the unfolded contraction;
//  * operands are plain row-major matrices (frames are materialised by fold_split_kernel): A box = 128 rows x
//    32 floats at row chain*M + m_tile*128, B box = 128 rows at row chain*n_bins_pad + n_tile*128;
//  * stage = A_hi, A_lo, B_hi, B_lo of 16 KB each = 64 KB -> 3 stages; MMAs are 128 x 128 x 8;
//  * M is the flattened frame index (tiles may straddle segments; the epilogue maps row -> (b, t));
//  * kF16 = true: operands are fp16 hi/lo planes (3xFP16: hi*hi + hi*lo + lo*hi, the same 22 operand bits as
//    3xTF32, at twice the MMA rate).  fp16 has tf32's mantissa but a 5-bit exponent, so the planes are
//    block-scaled by powers of two: every A row so that its largest element lies in [2^14, 2^15), the basis as
//    a whole likewise; elements whose lo part falls into the fp16 subnormals keep an ABSOLUTE error of 2^-25
//    in scaled units (2^-39 of the row maximum), far below the fp32 accumulation error.  The epilogue
//    multiplies by row_scale_inv[frame] * basis_scale_inv (exact).  A 128-byte swizzle row is 64 halves and
//    one MMA covers K = 16, so a stage carries twice the contraction length of the tf32 stage.
constexpr int F_BLOCK_N = 128;
constexpr int F_STAGES = 3;
constexpr int F_TILE_BYTES = 128 * BLOCK_K * 4;         // 16 KB (A and B tiles are both 128 rows)
constexpr int F_STAGE_BYTES = 4 * F_TILE_BYTES;         // 64 KB
constexpr int F_SMEM_BYTES = F_STAGES * F_STAGE_BYTES + BAR_BYTES + 1024;

struct FoldParams {
  int n_frames;               // frames per segment (T)
  int64_t m_rows;             // n_seg * n_frames
  int n_bins_pad, half;       // rows per basis plane, n_fft / 2
  int m_tiles, n_tiles;
  int epilogue, n_out_bins, n_store_bins;
  float power, w0;
  const float* p0;
  float* out0;
  const float* row_scale_inv;   // fp16 operands only: per-frame 2^-s_row
  float basis_scale_inv;        // fp16 operands only: 2^-s_basis
  // fused Mel projection (K1m, kernels instantiated with a MelTable): the epilogue does not store the spectrum
  float* mel_out;               // [n_seg][n_mels][n_frames], zeroed before the launch, accumulated with RED.ADD
  int n_mels;
  int corr_first;               // CTA-pair kernel: order of the three MMAs of the split product, see the kernel
};

// ---- epilogue of one unit (128 frames x 128 bins), shared by the one-CTA and the CTA-pair folded kernels ----
// Thread <-> frame row, 4 chunks of 32 bins.  Plain mode: format + store the spectrum.  Mel mode: the spectrum value
// P goes straight into the banded Mel projection.  The filterbank is pairwise overlapping (bin k feeds bands
// band0[k] and band0[k]+1, band0 non-decreasing), and k is warp-uniform, so two rotating register accumulators
// follow the band pair of the current bin; a finished band is added to mel_out[b][band][t] with one RED.ADD per
// thread -- lanes are consecutive frames, so a warp's RED covers 128 contiguous bytes.  A band receives at most two
// partial sums (one per tile it straddles) on top of the zero fill, so the result does not depend on their order.
// The loop over the tile's 128 bins is ROLLED, four bins per trip straight out of TMEM (tcgen05.ld ...x4): one
// epilogue warp runs alone on its scheduler, so straight-line code that overflows the instruction cache costs
// ~10 cycles per instruction (profiles/r01c: the unrolled variant was 40 % slower than the whole contraction).
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

struct MelAcc {
  int b0;
  float acc0, acc1;
  float* cur;        // &mel[b][b0][t]: advanced by one band row per rotation (no 64-bit multiply per flush)
};

// Eight consecutive bins of one frame into the rotating band accumulators.
// kFast (power spectrum, no rank-1 p0 term): the power-of-two operand scale is pulled out of the loop -- the
// accumulators hold sum w (re^2 + im^2) in scaled units and a finished band is multiplied by scale^2 once.  Scaling by
// a power of two commutes with every rounding, so the result is bit-identical to scaling each value.
// The Mel table travels as a 16 KB KERNEL PARAMETER (constant bank): the per-bin lookups are uniform constant
// loads that never touch shared memory -- the tensor pipe reads smem at its full 128 B/clk (64 wavefronts per
// 64-cycle MMA), so every LDS in the epilogue comes straight out of the contraction's operand bandwidth (with the
// table staged in smem the fused kernel was 218 us, profiles/r01e).
constexpr int kMelTableBins = 1024;
struct MelTable {
  float4 e[kMelTableBins];      // (w0, w1, band0 as int bits, -): bin k adds w0 P to band0, w1 P to band0 + 1
};
struct NoTable {
  int unused;
};

// (Tried out of line to shrink the loop: the call spills the accumulator registers around it, 304 us.)
__device__ __forceinline__ void mel_flush(float* __restrict__ dst, int band, int n_mels, float v, bool f_ok) {
  if (f_ok && v != 0.f && band < n_mels) atomicAdd(dst, v);
}

template <int kEpi, bool kFast>
__device__ __forceinline__ void mel_bins8(const FoldParams& p, const uint32_t (&re)[8], const uint32_t (&im)[8],
                                          const float4* tab, float* __restrict__ col, bool f_ok, float scale,
                                          float re_add, MelAcc& a) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 e = tab[i];                                       // constant bank, same index in every lane
    float v;
    if constexpr (kFast) {
      const float r = __uint_as_float(re[i]), m = __uint_as_float(im[i]);
      v = fmaf(r, r, m * m);
    } else {
      v = stft_value_t<kEpi>(p.power, fmaf(__uint_as_float(re[i]), scale, re_add), __uint_as_float(im[i]) * scale);
    }
    const int band = __float_as_int(e.z);
    // warp-uniform; NOT unrolled: nvcc otherwise emits four copies of the flush plus remainder logic per bin
    // (3 200 instructions per loop trip), which no longer fit the instruction cache the epilogue shares with the
    // MMA-issuing thread
#pragma unroll 1
    while (a.b0 < band) {
      mel_flush(a.cur, a.b0, p.n_mels, kFast ? a.acc0 * (scale * scale) : a.acc0, f_ok);
      a.cur += p.n_frames;
      a.acc0 = a.acc1;
      a.acc1 = 0.f;
      ++a.b0;
    }
    a.acc0 = fmaf(e.x, v, a.acc0);
    a.acc1 = fmaf(e.y, v, a.acc1);
  }
}

// The loop over the tile's 128 bins is ROLLED (eight bins per tcgen05.ld, two register sets so that the load of
// the next eight is in flight while these are folded in): one epilogue warp runs alone on its scheduler, so
// straight-line code that overflows the instruction cache costs ~10 cycles per instruction, and a TMEM load issued
// and awaited in the same trip costs its full latency 32 times per tile (profiles/r01d).
template <int kEpi, bool kFast>
__device__ __forceinline__ void mel_unit(const FoldParams& p, uint32_t taddr /* first column of the range */,
                                         const float4* tab /* first bin of the range */, int n_trips /* 8 bins each */,
                                         float* __restrict__ col, bool f_ok, float scale, float re_add) {
  MelAcc a{__float_as_int(tab[0].z), 0.f, 0.f, nullptr};
  a.cur = col + (int64_t)a.b0 * p.n_frames;
  uint32_t re0[8], im0[8], re1[8], im1[8];
  tmem_ld8(taddr, re0);
  tmem_ld8(taddr + 128, im0);
#pragma unroll 1
  for (int j = 0; j < n_trips; j += 2) {
    tmem_ld_wait();                                                // set 0 (bins 8j .. 8j+7) has landed
    tmem_ld8(taddr + 8 * (j + 1), re1);
    tmem_ld8(taddr + 128 + 8 * (j + 1), im1);
    mel_bins8<kEpi, kFast>(p, re0, im0, tab + 8 * j, col, f_ok, scale, re_add, a);
    tmem_ld_wait();                                                // set 1
    if (j + 2 < n_trips) {
      tmem_ld8(taddr + 8 * (j + 2), re0);
      tmem_ld8(taddr + 128 + 8 * (j + 2), im0);
    }
    mel_bins8<kEpi, kFast>(p, re1, im1, tab + 8 * (j + 1), col, f_ok, scale, re_add, a);
  }
  const float s2 = kFast ? scale * scale : 1.f;
  mel_flush(a.cur, a.b0, p.n_mels, a.acc0 * s2, f_ok);
  mel_flush(a.cur + p.n_frames, a.b0 + 1, p.n_mels, a.acc1 * s2, f_ok);
}

// c_begin, c_end: the 32-bin chunks of the tile this warp handles ([0, 4) = the whole tile; the CTA-pair kernel
// gives each of its two epilogue groups one half, which halves the epilogue's latency per tile).
template <typename TabT>
__device__ __forceinline__ void fold_epilogue_unit(const FoldParams& p, const TabT& tab, uint32_t taddr, bool f_ok,
                                                   int n_tile, int c_begin, int c_end, int b, int t, float scale,
                                                   float re0) {
  if constexpr (std::is_same<TabT, MelTable>::value) {
    const float4* tt = tab.e + n_tile * 128 + c_begin * 32;
    const uint32_t ta = taddr + (uint32_t)(c_begin * 32);
    const int trips = (c_end - c_begin) * 4;
    float* col = p.mel_out + (int64_t)b * p.n_mels * p.n_frames + t;
    if (p.epilogue == RVB_EPI_POWER && p.p0 == nullptr) mel_unit<RVB_EPI_POWER, true>(p, ta, tt, trips, col, f_ok, scale, 0.f);
    else if (p.epilogue == RVB_EPI_POWER) mel_unit<RVB_EPI_POWER, false>(p, ta, tt, trips, col, f_ok, scale, re0);
    else if (p.epilogue == RVB_EPI_MAGNITUDE) mel_unit<RVB_EPI_MAGNITUDE, false>(p, ta, tt, trips, col, f_ok, scale, re0);
    else mel_unit<RVB_EPI_POWER_P, false>(p, ta, tt, trips, col, f_ok, scale, re0);
    return;
  }
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    uint32_t re[32], im[32];
    tmem_ld32(taddr + c * 32, re);
    tmem_ld32(taddr + 128 + c * 32, im);
    tmem_ld_wait();
    const int k0 = n_tile * 128 + c * 32;
    if (f_ok)
      stft_store_chunk(p.epilogue, p.power, re, im, scale, re0, p.out0, b, k0, t, p.n_out_bins, p.n_store_bins,
                       p.n_frames);
  }
}

template <bool kF16>
__global__ void __launch_bounds__(NUM_THREADS, 1)
stft_gemm_fold_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                      const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                      const FoldParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + F_STAGES * F_STAGE_BYTES;
  auto s_tile = [&](int s, int which) { return smem_base + s * F_STAGE_BYTES + which * F_TILE_BYTES; };   // 0 A_hi 1 A_lo 2 B_hi 3 B_lo
  auto bar_full = [&](int s) { return bar_base + 8 * s; };
  auto bar_empty = [&](int s) { return bar_base + 8 * (F_STAGES + s); };
  auto bar_tmem_full = [&](int a) { return bar_base + 8 * (2 * F_STAGES + a); };
  auto bar_tmem_empty = [&](int a) { return bar_base + 8 * (2 * F_STAGES + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8 * (2 * F_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_a_lo);
    tma_prefetch_desc(&tm_b_hi);
    tma_prefetch_desc(&tm_b_lo);
    for (int s = 0; s < F_STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tmem_full(a), 1);
      mbar_init(bar_tmem_empty(a), 4);
    }
    fence_barrier_init();
  }
  __syncthreads();                                     // mbarrier words initialised before the allocator (another warp) writes beside them
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr) : "memory");

  constexpr int kBlockK = kF16 ? 2 * BLOCK_K : BLOCK_K;     // elements per 128-byte swizzle row
  const int n_units = p.m_tiles * p.n_tiles;
  const int kb_per_chain = p.half / kBlockK;
  const int num_kb = 2 * kb_per_chain;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int m_tile = unit / p.n_tiles, n_tile = unit - m_tile * p.n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int chain = kb >= kb_per_chain;
          const int kk = (kb - chain * kb_per_chain) * kBlockK;
          const int a_row = (int)(chain * p.m_rows) + m_tile * BLOCK_M;
          const int b_row = chain * p.n_bins_pad + n_tile * F_BLOCK_N;
          mbar_wait(bar_empty(stage), phase ^ 1u, nullptr, 1);
          mbar_expect_tx(bar_full(stage), F_STAGE_BYTES);
          tma_load_2d(&tm_a_hi, s_tile(stage, 0), bar_full(stage), kk, a_row);
          tma_load_2d(&tm_a_lo, s_tile(stage, 1), bar_full(stage), kk, a_row);
          tma_load_2d(&tm_b_hi, s_tile(stage, 2), bar_full(stage), kk, b_row);
          tma_load_2d(&tm_b_lo, s_tile(stage, 3), bar_full(stage), kk, b_row);
          if (++stage == F_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BLOCK_M, F_BLOCK_N, kF16 ? FMT_F16 : FMT_TF32);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        mbar_wait(bar_tmem_empty(acc), acc_phase ^ 1u, nullptr, 2);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          const int chain = kb >= kb_per_chain;
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS + chain * F_BLOCK_N);
          const bool first_kb = (kb == 0) || (kb == kb_per_chain);
          mbar_wait(bar_full(stage), phase, nullptr, 3);
          tc_fence_after();
          const uint64_t da_hi = make_sw128_desc(s_tile(stage, 0));
          const uint64_t da_lo = make_sw128_desc(s_tile(stage, 1));
          const uint64_t db_hi = make_sw128_desc(s_tile(stage, 2));
          const uint64_t db_lo = make_sw128_desc(s_tile(stage, 3));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)(k * UMMA_K * 4 >> 4);     // one MMA consumes 32 bytes of the row
            // small cross terms first: while the accumulator is still small their truncation costs nothing
            if constexpr (kF16) {
              umma_f16(d_tmem, da_hi + adv, db_lo + adv, idesc, (first_kb && k == 0) ? 0u : 1u);
              umma_f16(d_tmem, da_lo + adv, db_hi + adv, idesc, 1u);
              umma_f16(d_tmem, da_hi + adv, db_hi + adv, idesc, 1u);
            } else {
              umma_tf32(d_tmem, da_hi + adv, db_lo + adv, idesc, (first_kb && k == 0) ? 0u : 1u);
              umma_tf32(d_tmem, da_lo + adv, db_hi + adv, idesc, 1u);
              umma_tf32(d_tmem, da_hi + adv, db_hi + adv, idesc, 1u);
            }
          }
          umma_commit(bar_empty(stage));
          if (++stage == F_STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(bar_tmem_full(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      const int m_tile = unit / p.n_tiles, n_tile = unit - m_tile * p.n_tiles;
      const int64_t f = (int64_t)m_tile * BLOCK_M + row;        // flattened frame index
      const bool f_ok = f < p.m_rows;
      const int b = f_ok ? (int)(f / p.n_frames) : 0;
      const int t = f_ok ? (int)(f - (int64_t)b * p.n_frames) : 0;
      const float re0 = (p.p0 != nullptr && f_ok) ? p.w0 * __ldg(p.p0 + f) : 0.f;
      const float scale = (kF16 && f_ok) ? __ldg(p.row_scale_inv + f) * p.basis_scale_inv : 1.f;
      mbar_wait(bar_tmem_full(acc), acc_phase, nullptr, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_COLS);
      fold_epilogue_unit(p, NoTable{}, taddr, f_ok, n_tile, 0, 4, b, t, scale, re0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tmem_empty(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------- folded kernel, CTA pairs (K1f2)
// The one-CTA folded kernel moves 64 KB of operands per 12 MMAs and saturates the L2 -> SM path (13.6 TB/s over
// 148 SMs, profiles/r01c) at 55 % of the tensor pipe.  Here two CTAs of a cluster (one TPC) run
// tcgen05.mma.cta_group::2 on a 256-frame x 128-bin tile: each CTA stages its own 128 frame rows and only HALF of
// the basis tile (64 bins), the pair's tensor cores read both halves, so a CTA moves 48 KB per 12 MMAs and the
// ring is 4 stages deep.  Protocol (cf. the CUTLASS 2-SM kernels):
//   full[s]        lives in the leader (rank 0): the leader's producer arms it with the bytes of BOTH CTAs, and
//                  both CTAs' TMA loads (cp.async.bulk.tensor...cta_group::2) complete_tx on it;
//   empty[s]       one per CTA, released by the leader's tcgen05.commit...multicast::cluster (mask 0b11);
//   tmem_full[a]   one per CTA, same multicast commit;
//   tmem_empty[a]  lives in the leader, count 8: four epilogue warps of each CTA arrive (remotely from rank 1).
// Only the leader's warp 1 issues MMAs; each CTA's epilogue drains its own 128 TMEM lanes.
constexpr int P_STAGES = 4;
constexpr int P_EPI_GROUPS = 4;                       // epilogue groups of four warps; each takes 128 / P_EPI_GROUPS bins
constexpr int P_NUM_THREADS = 64 + P_EPI_GROUPS * 128;   // TMA warp, MMA warp, epilogue groups
constexpr int P_A_BYTES = 128 * 128;                    // 128 frame rows x one 128-byte swizzle row
constexpr int P_B_BYTES = 64 * 128;                     // this CTA's half of the 128 basis rows
constexpr int P_STAGE_BYTES = 2 * P_A_BYTES + 2 * P_B_BYTES;     // A_hi A_lo B_hi B_lo = 48 KB
constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + BAR_BYTES + 1024;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* tm, uint32_t smem_dst, uint32_t leader_bar, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}

template <typename TabT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_NUM_THREADS, 1)
stft_gemm_fold_pair_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                           const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                           const FoldParams p, const __grid_constant__ TabT tab) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + P_STAGES * P_STAGE_BYTES;
  auto s_a = [&](int s, int lo) { return smem_base + s * P_STAGE_BYTES + lo * P_A_BYTES; };
  auto s_b = [&](int s, int lo) { return smem_base + s * P_STAGE_BYTES + 2 * P_A_BYTES + lo * P_B_BYTES; };
  auto bar_full = [&](int s) { return bar_base + 8 * s; };
  auto bar_empty = [&](int s) { return bar_base + 8 * (P_STAGES + s); };
  auto bar_tmem_full = [&](int a) { return bar_base + 8 * (2 * P_STAGES + a); };
  auto bar_tmem_empty = [&](int a) { return bar_base + 8 * (2 * P_STAGES + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8 * (2 * P_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_a_lo);
    tma_prefetch_desc(&tm_b_hi);
    tma_prefetch_desc(&tm_b_lo);
    for (int s = 0; s < P_STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tmem_full(a), 1);
      mbar_init(bar_tmem_empty(a), 8 * P_EPI_GROUPS);   // epilogue warps of both CTAs (only the leader's copy is used)
    }
    fence_barrier_init();
  }
  __syncthreads();                                     // mbarrier words initialised before the allocator (another warp) writes beside them
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                             // both CTAs' barriers are initialised before any remote signal
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr) : "memory");

  constexpr int kBlockK = 2 * BLOCK_K;            // 64 halves per 128-byte swizzle row
  const int n_units = p.m_tiles * p.n_tiles;      // m_tiles counts 256-frame pair tiles here
  const int kb_per_chain = p.half / kBlockK;
  // Order of the split product's three MMAs.  The tensor core adds every MMA into the fp32 accumulator with
  // TRUNCATION, not round-to-nearest: each of the 3 * half / 16 additions of a chain loses up to an ulp of the running
  // sum, always in the same direction, so a weak bin whose partial sums carry a strong partial's energy ends up biased
  // by ~(number of additions) * ulp(partial sum) / 2 (profiles/r02_precision.md).  corr_first: the chain runs twice --
  // first hi*lo + lo*hi over the whole contraction (partial sums 2^-11 of the product's: their truncation is
  // negligible), then hi*hi on top -- one truncation at full magnitude per 16 terms instead of three.  The second
  // pass streams the hi planes again: 1.5x the operand traffic, the same MMAs.
  const int steps_per_chain = p.corr_first ? 2 * kb_per_chain : kb_per_chain;
  const int n_steps = 2 * steps_per_chain;
  const int unit0 = (int)(blockIdx.x >> 1), unit_step = (int)(gridDim.x >> 1);     // cluster id / count (cluster = 2 CTAs)

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t leader_full0 = map_to_rank(bar_full(0), 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = unit0; unit < n_units; unit += unit_step) {
        const int m_tile = unit / p.n_tiles, n_tile = unit - m_tile * p.n_tiles;
        for (int step = 0; step < n_steps; ++step) {
          const int chain = step >= steps_per_chain;
          const int in_chain = step - chain * steps_per_chain;
          const bool hi_only = in_chain >= kb_per_chain;          // second pass of a corrections-first chain
          const int kk = (in_chain - (hi_only ? kb_per_chain : 0)) * kBlockK;
          const int a_row = (int)(chain * p.m_rows) + m_tile * 256 + (int)rank * 128;
          const int b_row = chain * p.n_bins_pad + n_tile * F_BLOCK_N + (int)rank * 64;
          mbar_wait(bar_empty(stage), phase ^ 1u, nullptr, 1);
          if (leader) mbar_expect_tx(bar_full(stage), hi_only ? 2 * (P_A_BYTES + P_B_BYTES) : 2 * P_STAGE_BYTES);
          const uint32_t fb = leader_full0 + 8 * stage;
          tma_load_2d_pair(&tm_a_hi, s_a(stage, 0), fb, kk, a_row);
          tma_load_2d_pair(&tm_b_hi, s_b(stage, 0), fb, kk, b_row);
          if (!hi_only) {
            tma_load_2d_pair(&tm_a_lo, s_a(stage, 1), fb, kk, a_row);
            tma_load_2d_pair(&tm_b_lo, s_b(stage, 1), fb, kk, b_row);
          }
          if (++stage == P_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc(256, F_BLOCK_N, FMT_F16);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int unit = unit0; unit < n_units; unit += unit_step) {
        mbar_wait(bar_tmem_empty(acc), acc_phase ^ 1u, nullptr, 2);
        tc_fence_after();
        for (int step = 0; step < n_steps; ++step) {
          const int chain = step >= steps_per_chain;
          const int in_chain = step - chain * steps_per_chain;
          const bool hi_only = in_chain >= kb_per_chain;
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS + chain * F_BLOCK_N);
          const bool first_kb = in_chain == 0;
          mbar_wait(bar_full(stage), phase, nullptr, 3);
          tc_fence_after();
          const uint64_t da_hi = make_sw128_desc(s_a(stage, 0));
          const uint64_t da_lo = make_sw128_desc(s_a(stage, 1));
          const uint64_t db_hi = make_sw128_desc(s_b(stage, 0));
          const uint64_t db_lo = make_sw128_desc(s_b(stage, 1));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);             // one MMA consumes 32 bytes of the row
            if (!hi_only) {
              umma_f16_pair(d_tmem, da_hi + adv, db_lo + adv, idesc, (first_kb && k == 0) ? 0u : 1u);
              umma_f16_pair(d_tmem, da_lo + adv, db_hi + adv, idesc, 1u);
            }
            if (hi_only || !p.corr_first) umma_f16_pair(d_tmem, da_hi + adv, db_hi + adv, idesc, 1u);
          }
          umma_commit_pair(bar_empty(stage));       // frees the slot in both CTAs
          if (++stage == P_STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit_pair(bar_tmem_full(acc));       // accumulators complete in both CTAs
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // P_EPI_GROUPS x 4 epilogue warps per CTA: group g folds bins [32 g, 32 g + 32) of every tile (TMEM lane quarter
    // = warp % 4 in every group).  A TMEM accumulator stage is handed back to the tensor pipe only when its epilogue
    // has finished, so what must stay below one MMA tile time is the epilogue's LATENCY per tile, whatever its
    // throughput: one group 222 us, two 197 us, four 190 us = the contraction without any Mel work (profiles/r01e).
    // A band straddling a 32-bin boundary gets its two partial sums from two groups (or two tiles): still at most
    // two per element for banks whose bands are narrower than 33 bins -- checked on the host.
    const int half_id = (warp - EPI_WARP0) >> 2;
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t leader_tmem_empty0 = map_to_rank(bar_tmem_empty(0), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = unit0; unit < n_units; unit += unit_step) {
      const int m_tile = unit / p.n_tiles, n_tile = unit - m_tile * p.n_tiles;
      const int64_t f = (int64_t)m_tile * 256 + (int64_t)rank * 128 + row;       // flattened frame index
      const bool f_ok = f < p.m_rows;
      const int b = f_ok ? (int)(f / p.n_frames) : 0;
      const int t = f_ok ? (int)(f - (int64_t)b * p.n_frames) : 0;
      const float re0 = (p.p0 != nullptr && f_ok) ? p.w0 * __ldg(p.p0 + f) : 0.f;
      const float scale = f_ok ? __ldg(p.row_scale_inv + f) * p.basis_scale_inv : 1.f;
      mbar_wait(bar_tmem_full(acc), acc_phase, nullptr, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_COLS);
      constexpr int kChunks = 4 / P_EPI_GROUPS;
      fold_epilogue_unit(p, tab, taddr, f_ok, n_tile, kChunks * half_id, kChunks * half_id + kChunks, b, t, scale, re0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tmem_empty0 + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                             // the peer may still be reading our smem / signalling our barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------- twice-folded kernel, CTA pairs (K1q)
// One more symmetry than K1f2.  For integer bins, cos(2 pi (N/2-k) n / N) = (-1)^n cos(2 pi k n / N) and
// sin(2 pi (N/2-k) n / N) = -(-1)^n sin(2 pi k n / N), whatever the window.  Splitting the folded sums by the parity of n,
//     Ce = sum_{n even} Bc[k][n] e[n],  Co = sum_{n odd} Bc[k][n] e[n],  Se, So likewise with Bs and o,
// gives BOTH bin k (re = Ce + Co, im = Se + So) and bin N/2 - k (re = Ce - Co, im = -(Se - So)): one radix-2
// decimation step of the FFT.  The contraction runs over k = 1 .. N/4 only: four chains of length N/4 per tile
// instead of two of length N/2 over twice as many rows -- HALF the multiply-adds of K1f2 for the same 1022 bins (bins 0
// and N/2 are not produced; the Mel bank does not read them -- checked on the host).
//  * A planes: [hi|lo][e|o][frame][N/2] with the even-n columns first (rvb_fold_split2_f16); chain c reads plane
//    c >> 1 at column offset (c & 1) N/4.  B planes: [hi|lo][Ce|Co|Se|So][N/4 rows][N/4], row r <-> k = r + 1.
//  * tile = 256 frames x 64 k-values; MMAs are M256 N64 K16 (cta_group::2), each CTA stages 128 frame rows and 32
//    basis rows per 64-column K-block: 40 KB per stage, 5 stages.  TMEM: 4 accumulators x 64 columns per stage, 2 stages.
//  * epilogue: four groups of four warps.  Group g = (stream g >> 1, half g & 1) walks 32 of the tile's k-values in
//    ascending order and feeds P = (Ce +- Co)^2 + (Se +- So)^2 to the rotating band accumulators.  Stream 0 is the
//    ascending bins k; stream 1 is bins N/2 - k, DEscending in frequency -- in reversed band coordinates
//    (band' = n_mels - 1 - band) that walk is ascending again, so both streams run the SAME code with a different sign,
//    table half and output stride (basis.mel_epilogue_table2).  A band receives at most two partial sums (32-row
//    chunks; checked on the host), so the result stays bit-reproducible.
constexpr int Q_BLOCK_N = 64;
constexpr int Q_STAGES = 5;
constexpr int Q_B_BYTES = 32 * 128;                     // this CTA's half of the 64 basis rows
constexpr int Q_STAGE_BYTES = 2 * P_A_BYTES + 2 * Q_B_BYTES;     // A_hi A_lo B_hi B_lo = 40 KB
constexpr int Q_SMEM_BYTES = Q_STAGES * Q_STAGE_BYTES + BAR_BYTES + 1024;
constexpr int Q_CHUNK = 32;                             // k-values per epilogue group and tile

struct Fold2Params {
  int n_frames;               // frames per segment (T)
  int64_t m_rows;             // n_seg * n_frames
  int n_k, quarter;           // basis rows per chain (N/4), contraction length per chain (N/4)
  int m_tiles, n_tiles;
  const float* row_scale_inv;
  float basis_scale_inv;
  float* mel_out;             // [2][n_seg][n_mels][n_frames] (cos^2 part | sin^2 part), zeroed before the launch
  int64_t plane_stride;       // n_seg * n_mels * n_frames
  int n_mels;
  int exp;                    // ablation builds only (make ABLATION=1): RVB_EXP bits 1: no RED.ADD, 2: no epilogue work,
                              // 4: relaxed hand-back; always 0 in the product build
};

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}

struct Mel2Acc {
  int b0;
  float acc0, acc1;
  float* cur;        // row of band b0 (stream 1: of band n_mels - 1 - b0); advanced by `stride` per rotation
};

// Four consecutive k-values of one frame.  (ce, co, se, so): the four accumulators; sgn = +1 / -1 selects bin k / N/2-k.
__device__ __forceinline__ void mel2_bins4(const uint32_t (&ce)[4], const uint32_t (&co)[4], const uint32_t (&se)[4],
                                           const uint32_t (&so)[4], const float4* tab, float sgn, int n_mels,
                                           int64_t stride, bool f_ok, float scale2, Mel2Acc& a) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 e = tab[i];                                       // constant bank, same index in every lane
    const float r = fmaf(sgn, __uint_as_float(co[i]), __uint_as_float(ce[i]));
    const float m = fmaf(sgn, __uint_as_float(so[i]), __uint_as_float(se[i]));
    const float v = fmaf(r, r, m * m);
    const int band = __float_as_int(e.z);
#pragma unroll 1
    while (a.b0 < band) {                                          // warp-uniform; rolled (instruction cache, see K1m)
      const float out = a.acc0 * scale2;
      if (f_ok && out != 0.f && (unsigned)a.b0 < (unsigned)n_mels) atomicAdd(a.cur, out);
      a.cur += stride;
      a.acc0 = a.acc1;
      a.acc1 = 0.f;
      ++a.b0;
    }
    a.acc0 = fmaf(e.x, v, a.acc0);
    a.acc1 = fmaf(e.y, v, a.acc1);
  }
}

__device__ __forceinline__ void mel2_unit(const Fold2Params& p, uint32_t taddr /* column of the first k-value in Ce */,
                                          const float4* tab, int stream, float* __restrict__ col, bool f_ok, float scale) {
  const float sgn = stream ? -1.f : 1.f;
  const int64_t stride = stream ? -(int64_t)p.n_frames : (int64_t)p.n_frames;
  Mel2Acc a{__float_as_int(tab[0].z), 0.f, 0.f, nullptr};
  a.cur = col + (int64_t)(stream ? p.n_mels - 1 - a.b0 : a.b0) * p.n_frames;
  const float scale2 = scale * scale;
  uint32_t ce0[4], co0[4], se0[4], so0[4], ce1[4], co1[4], se1[4], so1[4];
  tmem_ld4(taddr, ce0);
  tmem_ld4(taddr + Q_BLOCK_N, co0);
  tmem_ld4(taddr + 2 * Q_BLOCK_N, se0);
  tmem_ld4(taddr + 3 * Q_BLOCK_N, so0);
#pragma unroll 1
  for (int j = 0; j < Q_CHUNK / 4; j += 2) {
    tmem_ld_wait();                                                // set 0 (k-values 4j .. 4j+3) has landed
    tmem_ld4(taddr + 4 * (j + 1), ce1);
    tmem_ld4(taddr + Q_BLOCK_N + 4 * (j + 1), co1);
    tmem_ld4(taddr + 2 * Q_BLOCK_N + 4 * (j + 1), se1);
    tmem_ld4(taddr + 3 * Q_BLOCK_N + 4 * (j + 1), so1);
    mel2_bins4(ce0, co0, se0, so0, tab + 4 * j, sgn, p.n_mels, stride, f_ok, scale2, a);
    tmem_ld_wait();                                                // set 1
    if (j + 2 < Q_CHUNK / 4) {
      tmem_ld4(taddr + 4 * (j + 2), ce0);
      tmem_ld4(taddr + Q_BLOCK_N + 4 * (j + 2), co0);
      tmem_ld4(taddr + 2 * Q_BLOCK_N + 4 * (j + 2), se0);
      tmem_ld4(taddr + 3 * Q_BLOCK_N + 4 * (j + 2), so0);
    }
    mel2_bins4(ce1, co1, se1, so1, tab + 4 * (j + 1), sgn, p.n_mels, stride, f_ok, scale2, a);
  }
  float out = a.acc0 * scale2;
  if (f_ok && out != 0.f && (unsigned)a.b0 < (unsigned)p.n_mels) atomicAdd(a.cur, out);
  out = a.acc1 * scale2;
  if (f_ok && out != 0.f && (unsigned)(a.b0 + 1) < (unsigned)p.n_mels) atomicAdd(a.cur + stride, out);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_NUM_THREADS, 1)
stft_gemm_fold2_pair_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                            const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                            const Fold2Params p, const __grid_constant__ MelTable tab) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Q_STAGES * Q_STAGE_BYTES;
  auto s_a = [&](int s, int lo) { return smem_base + s * Q_STAGE_BYTES + lo * P_A_BYTES; };
  auto s_b = [&](int s, int lo) { return smem_base + s * Q_STAGE_BYTES + 2 * P_A_BYTES + lo * Q_B_BYTES; };
  auto bar_full = [&](int s) { return bar_base + 8 * s; };
  auto bar_empty = [&](int s) { return bar_base + 8 * (Q_STAGES + s); };
  auto bar_tmem_full = [&](int a) { return bar_base + 8 * (2 * Q_STAGES + a); };
  auto bar_tmem_empty = [&](int a) { return bar_base + 8 * (2 * Q_STAGES + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8 * (2 * Q_STAGES + 4);
  static_assert(8 * (2 * Q_STAGES + 4) + 4 <= BAR_BYTES, "barrier block too small");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_a_lo);
    tma_prefetch_desc(&tm_b_hi);
    tma_prefetch_desc(&tm_b_lo);
    for (int s = 0; s < Q_STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tmem_full(a), 1);
      mbar_init(bar_tmem_empty(a), 8 * P_EPI_GROUPS);   // epilogue warps of both CTAs (only the leader's copy is used)
    }
    fence_barrier_init();
  }
  __syncthreads();                                     // mbarrier words initialised before the allocator (another warp) writes beside them
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr) : "memory");

  constexpr int kBlockK = 2 * BLOCK_K;            // 64 halves per 128-byte swizzle row
  const int n_units = p.m_tiles * p.n_tiles;      // 256-frame x 64-k tiles
  const int kb_per_chain = p.quarter / kBlockK;
  const int num_kb = 4 * kb_per_chain;
  const int unit0 = (int)(blockIdx.x >> 1), unit_step = (int)(gridDim.x >> 1);

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t leader_full0 = map_to_rank(bar_full(0), 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = unit0; unit < n_units; unit += unit_step) {
        const int m_tile = unit / p.n_tiles, n_tile = unit - m_tile * p.n_tiles;
        int chain = 0, kbi = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int a_row = (int)((chain >> 1) * p.m_rows) + m_tile * 256 + (int)rank * 128;
          const int a_col = (chain & 1) * p.quarter + kbi * kBlockK;
          const int b_row = chain * p.n_k + n_tile * Q_BLOCK_N + (int)rank * 32;
          mbar_wait(bar_empty(stage), phase ^ 1u, nullptr, 1);
          if (leader) mbar_expect_tx(bar_full(stage), 2 * Q_STAGE_BYTES);
          const uint32_t fb = leader_full0 + 8 * stage;
          tma_load_2d_pair(&tm_a_hi, s_a(stage, 0), fb, a_col, a_row);
          tma_load_2d_pair(&tm_a_lo, s_a(stage, 1), fb, a_col, a_row);
          tma_load_2d_pair(&tm_b_hi, s_b(stage, 0), fb, kbi * kBlockK, b_row);
          tma_load_2d_pair(&tm_b_lo, s_b(stage, 1), fb, kbi * kBlockK, b_row);
          if (++stage == Q_STAGES) { stage = 0; phase ^= 1u; }
          if (++kbi == kb_per_chain) { kbi = 0; ++chain; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc(256, Q_BLOCK_N, FMT_F16);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int unit = unit0; unit < n_units; unit += unit_step) {
        mbar_wait(bar_tmem_empty(acc), acc_phase ^ 1u, nullptr, 2);
        tc_fence_after();
        int chain = 0, kbi = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS + chain * Q_BLOCK_N);
          const bool first_kb = kbi == 0;
          mbar_wait(bar_full(stage), phase, nullptr, 3);
          tc_fence_after();
          const uint64_t da_hi = make_sw128_desc(s_a(stage, 0));
          const uint64_t da_lo = make_sw128_desc(s_a(stage, 1));
          const uint64_t db_hi = make_sw128_desc(s_b(stage, 0));
          const uint64_t db_lo = make_sw128_desc(s_b(stage, 1));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);             // one MMA consumes 32 bytes of the row
            umma_f16_pair(d_tmem, da_hi + adv, db_lo + adv, idesc, (first_kb && k == 0) ? 0u : 1u);
            umma_f16_pair(d_tmem, da_lo + adv, db_hi + adv, idesc, 1u);
            umma_f16_pair(d_tmem, da_hi + adv, db_hi + adv, idesc, 1u);
          }
          umma_commit_pair(bar_empty(stage));
          if (++stage == Q_STAGES) { stage = 0; phase ^= 1u; }
          if (++kbi == kb_per_chain) { kbi = 0; ++chain; }
        }
        umma_commit_pair(bar_tmem_full(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    const int group = (warp - EPI_WARP0) >> 2;      // 0 .. 3
    const int stream = group >> 1, hhalf = group & 1;
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t leader_tmem_empty0 = map_to_rank(bar_tmem_empty(0), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = unit0; unit < n_units; unit += unit_step) {
      const int m_tile = unit / p.n_tiles, n_tile = unit - m_tile * p.n_tiles;
      const int64_t f = (int64_t)m_tile * 256 + (int64_t)rank * 128 + row;       // flattened frame index
      const bool f_ok = f < p.m_rows;
      const int b = f_ok ? (int)(f / p.n_frames) : 0;
      const int t = f_ok ? (int)(f - (int64_t)b * p.n_frames) : 0;
      const float scale = f_ok ? __ldg(p.row_scale_inv + f) * p.basis_scale_inv : 1.f;
      float* col = p.mel_out + (int64_t)b * p.n_mels * p.n_frames + t;
      const float4* tt = tab.e + stream * p.n_k + n_tile * Q_BLOCK_N + hhalf * Q_CHUNK;
      mbar_wait(bar_tmem_full(acc), acc_phase, nullptr, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_COLS + hhalf * Q_CHUNK);
      mel2_unit(p, taddr, tt, stream, col, f_ok, scale);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tmem_empty0 + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------- twice-folded kernel, one component per unit (K1q)
// stft_gemm_fold2_pair_kernel above moves 40 KB of operands per CTA for twelve M256 N64 MMAs: at the tensor pipe's pace
// that is 17 TB/s of L2 -> SM traffic over the chip, more than the L2 delivers (~12 TB/s): it runs L2-bound.  The Mel
// projection is LINEAR in the power P = re^2 + im^2, so the two components never have to meet in one thread: this
// kernel's unit is (256 frames, 128 k-values, ONE component): two chains (even n, odd n) of the e plane against the
// cos rows -- or of the o plane against the sin rows -- into two 128-column accumulators, exactly the shape of K1f2
// (48 KB per twelve M256 N128 MMAs, 4 stages) at half its contraction length.  The epilogue forms r = E +- O, adds
// w r^2 to the band accumulators and flushes them into the component's OWN Mel plane (mel_out[comp]); a plane
// element receives at most two partial sums, so both planes -- and their sum, taken by the consumer -- are
// bit-reproducible.  L2 -> SM traffic: 0.98 GB per 32 segments instead of 1.64 GB.
// scale2 == 0 for rows past the end (their accumulators are finite: the flush then adds nothing).
__device__ __forceinline__ void mel2c_bins8(const uint32_t (&ev)[8], const uint32_t (&od)[8], const float4* tab, float sgn,
                                            int n_mels, int64_t stride, float scale2, Mel2Acc& a) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 e = tab[i];                                       // constant bank, same index in every lane
    const float r = fmaf(sgn, __uint_as_float(od[i]), __uint_as_float(ev[i]));
    const float v = r * r;
    const int band = __float_as_int(e.z);
#pragma unroll 1
    while (a.b0 < band) {                                          // warp-uniform; rolled (instruction cache, see K1m)
      const float out = a.acc0 * scale2;
      if (out != 0.f && (unsigned)a.b0 < (unsigned)n_mels) atomicAdd(a.cur, out);
      a.cur += stride;
      a.acc0 = a.acc1;
      a.acc1 = 0.f;
      ++a.b0;
    }
    a.acc0 = fmaf(e.x, v, a.acc0);
    a.acc1 = fmaf(e.y, v, a.acc1);
  }
}

constexpr int C_CHUNK = 64;                             // k-values per epilogue group and tile

__device__ __forceinline__ void mel2c_unit(const Fold2Params& p, uint32_t taddr /* column of the first k-value in E */,
                                           const float4* tab, int stream, float* __restrict__ col, bool f_ok, float scale) {
  const float sgn = stream ? -1.f : 1.f;
  int n_mels = p.n_mels;
  asm volatile("" : "+r"(n_mels));                                 // keep it in a register (not an LDC per flush)
  const int64_t stride = stream ? -(int64_t)p.n_frames : (int64_t)p.n_frames;
  Mel2Acc a{__float_as_int(tab[0].z), 0.f, 0.f, nullptr};
  a.cur = col + (int64_t)(stream ? n_mels - 1 - a.b0 : a.b0) * p.n_frames;
  const float scale2 = f_ok ? scale * scale : 0.f;
  uint32_t e0[8], o0[8], e1[8], o1[8];
  tmem_ld8(taddr, e0);
  tmem_ld8(taddr + F_BLOCK_N, o0);
#pragma unroll 1
  for (int j = 0; j < C_CHUNK / 8; j += 2) {
    tmem_ld_wait();                                                // set 0 (k-values 8j .. 8j+7) has landed
    tmem_ld8(taddr + 8 * (j + 1), e1);
    tmem_ld8(taddr + F_BLOCK_N + 8 * (j + 1), o1);
    mel2c_bins8(e0, o0, tab + 8 * j, sgn, n_mels, stride, scale2, a);
    tmem_ld_wait();                                                // set 1
    if (j + 2 < C_CHUNK / 8) {
      tmem_ld8(taddr + 8 * (j + 2), e0);
      tmem_ld8(taddr + F_BLOCK_N + 8 * (j + 2), o0);
    }
    mel2c_bins8(e1, o1, tab + 8 * (j + 1), sgn, n_mels, stride, scale2, a);
  }
  float out = a.acc0 * scale2;
  if (out != 0.f && (unsigned)a.b0 < (unsigned)n_mels) atomicAdd(a.cur, out);
  out = a.acc1 * scale2;
  if (out != 0.f && (unsigned)(a.b0 + 1) < (unsigned)n_mels) atomicAdd(a.cur + stride, out);
}

template <int kStages>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_NUM_THREADS, 1)
stft_gemm_fold2c_pair_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                             const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                             const Fold2Params p, const __grid_constant__ MelTable tab) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * P_STAGE_BYTES;
  auto s_a = [&](int s, int lo) { return smem_base + s * P_STAGE_BYTES + lo * P_A_BYTES; };
  auto s_b = [&](int s, int lo) { return smem_base + s * P_STAGE_BYTES + 2 * P_A_BYTES + lo * P_B_BYTES; };
  auto bar_full = [&](int s) { return bar_base + 8 * s; };
  auto bar_empty = [&](int s) { return bar_base + 8 * (kStages + s); };
  auto bar_tmem_full = [&](int a) { return bar_base + 8 * (2 * kStages + a); };
  auto bar_tmem_empty = [&](int a) { return bar_base + 8 * (2 * kStages + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8 * (2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_a_lo);
    tma_prefetch_desc(&tm_b_hi);
    tma_prefetch_desc(&tm_b_lo);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tmem_full(a), 1);
      mbar_init(bar_tmem_empty(a), 8 * P_EPI_GROUPS);
    }
    fence_barrier_init();
  }
  __syncthreads();                                     // mbarrier words initialised before the allocator (another warp) writes beside them
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr) : "memory");

  constexpr int kBlockK = 2 * BLOCK_K;            // 64 halves per 128-byte swizzle row
  const int units_per_m = 2 * p.n_tiles;          // (component, 128-k tile)
  const int n_units = p.m_tiles * units_per_m;
  const int kb_per_chain = p.quarter / kBlockK;
  const int num_kb = 2 * kb_per_chain;            // even-n chain, odd-n chain
  const int unit0 = (int)(blockIdx.x >> 1), unit_step = (int)(gridDim.x >> 1);

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t leader_full0 = map_to_rank(bar_full(0), 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = unit0; unit < n_units; unit += unit_step) {
        const int m_tile = unit / units_per_m, rest = unit - m_tile * units_per_m;
        const int comp = rest / p.n_tiles, n_tile = rest - comp * p.n_tiles;
        const int a_row = (int)(comp * p.m_rows) + m_tile * 256 + (int)rank * 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int chain = kb >= kb_per_chain;
          const int kk = (kb - chain * kb_per_chain) * kBlockK;
          const int b_row = (2 * comp + chain) * p.n_k + n_tile * F_BLOCK_N + (int)rank * 64;
          mbar_wait(bar_empty(stage), phase ^ 1u, nullptr, 1);
          if (leader) mbar_expect_tx(bar_full(stage), 2 * P_STAGE_BYTES);
          const uint32_t fb = leader_full0 + 8 * stage;
          tma_load_2d_pair(&tm_a_hi, s_a(stage, 0), fb, chain * p.quarter + kk, a_row);
          tma_load_2d_pair(&tm_a_lo, s_a(stage, 1), fb, chain * p.quarter + kk, a_row);
          tma_load_2d_pair(&tm_b_hi, s_b(stage, 0), fb, kk, b_row);
          tma_load_2d_pair(&tm_b_lo, s_b(stage, 1), fb, kk, b_row);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc(256, F_BLOCK_N, FMT_F16);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int unit = unit0; unit < n_units; unit += unit_step) {
        mbar_wait(bar_tmem_empty(acc), acc_phase ^ 1u, nullptr, 2);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          const int chain = kb >= kb_per_chain;
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS + chain * F_BLOCK_N);
          const bool first_kb = (kb == 0) || (kb == kb_per_chain);
          mbar_wait(bar_full(stage), phase, nullptr, 3);
          tc_fence_after();
          const uint64_t da_hi = make_sw128_desc(s_a(stage, 0));
          const uint64_t da_lo = make_sw128_desc(s_a(stage, 1));
          const uint64_t db_hi = make_sw128_desc(s_b(stage, 0));
          const uint64_t db_lo = make_sw128_desc(s_b(stage, 1));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);             // one MMA consumes 32 bytes of the row
            umma_f16_pair(d_tmem, da_hi + adv, db_lo + adv, idesc, (first_kb && k == 0) ? 0u : 1u);
            umma_f16_pair(d_tmem, da_lo + adv, db_hi + adv, idesc, 1u);
            umma_f16_pair(d_tmem, da_hi + adv, db_hi + adv, idesc, 1u);
          }
          umma_commit_pair(bar_empty(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_pair(bar_tmem_full(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    const int group = (warp - EPI_WARP0) >> 2;      // 0 .. 3
    const int stream = group >> 1, hhalf = group & 1;
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t leader_tmem_empty0 = map_to_rank(bar_tmem_empty(0), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = unit0; unit < n_units; unit += unit_step) {
      const int m_tile = unit / units_per_m, rest = unit - m_tile * units_per_m;
      const int comp = rest / p.n_tiles, n_tile = rest - comp * p.n_tiles;
      const int64_t f = (int64_t)m_tile * 256 + (int64_t)rank * 128 + row;       // flattened frame index
      const bool f_ok = f < p.m_rows;
      const int b = f_ok ? (int)(f / p.n_frames) : 0;
      const int t = f_ok ? (int)(f - (int64_t)b * p.n_frames) : 0;
      const float scale = f_ok ? __ldg(p.row_scale_inv + f) * p.basis_scale_inv : 1.f;
      float* col = p.mel_out + comp * p.plane_stride + (int64_t)b * p.n_mels * p.n_frames + t;
      const float4* tt = tab.e + stream * p.n_k + n_tile * F_BLOCK_N + hhalf * C_CHUNK;
      mbar_wait(bar_tmem_full(acc), acc_phase, nullptr, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_COLS + hhalf * C_CHUNK);
#ifdef RVB_ABLATION   // measurement-only build (make ABLATION=1): RVB_EXP switches parts of the epilogue off -- wrong results
      if (!(p.exp & 2)) mel2c_unit(p, taddr, tt, stream, col, f_ok && !(p.exp & 1), scale);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (p.exp & 4) mbar_arrive_cluster_relaxed(leader_tmem_empty0 + 8 * acc);
        else mbar_arrive_cluster(leader_tmem_empty0 + 8 * acc);
      }
#else
      mel2c_unit(p, taddr, tt, stream, col, f_ok, scale);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tmem_empty0 + 8 * acc);
#endif
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------- K1q with the frame tiles shared across a cluster (K1qm)
// stft_gemm_fold2c_pair_kernel reads every 128-row frame tile from the L2 once per (k tile): four times per component.
// Here a cluster holds kPairs CTA pairs that work on the SAME 256 frames and component and on kPairs DIFFERENT k tiles:
// rank = 2 * pair + half.  The frame tile of the CTAs with the same `half` is the same, so each of them fetches
// 128 / kPairs of its rows and TMA-multicasts them to all kPairs CTAs of that half: the L2 -> SM traffic of the frame
// operand drops by kPairs (48 KB -> 16 + 32 / kPairs KB per CTA and stage).  Protocol changes against K1q:
//   full[s]   (leader of each pair) still one arrive.expect_tx for the pair's 96 KB; the bytes now come from the
//             pair's own basis loads and from the multicast frame slices of ALL pairs (complete_tx lands on the leader
//             of every destination pair: mbarrier address with the peer bit cleared, relative to the destination CTA)
//   empty[s]  (every CTA) kPairs arrivals: the MMA thread of EVERY pair commits to the barrier of EVERY CTA of the
//             cluster, because any CTA's producer writes into every same-half CTA's stage s
//   tmem_*    per pair, as in K1q
__device__ __forceinline__ void tma_load_2d_pair_mc(const CUtensorMap* tm, uint32_t smem_dst, uint32_t bar_local, int c0,
                                                    int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_local & 0xFEFFFFFFu), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mask(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(mask)
      : "memory");
}

template <int kStages, int kPairs>
__global__ void __launch_bounds__(P_NUM_THREADS, 1)
stft_gemm_fold2c_mc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                           const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                           const Fold2Params p, const __grid_constant__ MelTable tab) {
  constexpr int kSliceRows = 128 / kPairs;
  constexpr int kSliceBytes = kSliceRows * 128;
  static_assert(kSliceBytes % 1024 == 0, "a slice must be whole swizzle atoms");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * P_STAGE_BYTES;
  auto s_a = [&](int s, int lo) { return smem_base + s * P_STAGE_BYTES + lo * P_A_BYTES; };
  auto s_b = [&](int s, int lo) { return smem_base + s * P_STAGE_BYTES + 2 * P_A_BYTES + lo * P_B_BYTES; };
  auto bar_full = [&](int s) { return bar_base + 8 * s; };
  auto bar_empty = [&](int s) { return bar_base + 8 * (kStages + s); };
  auto bar_tmem_full = [&](int a) { return bar_base + 8 * (2 * kStages + a); };
  auto bar_tmem_empty = [&](int a) { return bar_base + 8 * (2 * kStages + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8 * (2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t half_id = rank & 1u, pair = rank >> 1;
  const bool leader = half_id == 0;
  const uint32_t leader_rank = rank & ~1u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_a_lo);
    tma_prefetch_desc(&tm_b_hi);
    tma_prefetch_desc(&tm_b_lo);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), kPairs);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tmem_full(a), 1);
      mbar_init(bar_tmem_empty(a), 8 * P_EPI_GROUPS);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                             // every CTA's barriers are initialised before any remote signal
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr) : "memory");

  constexpr int kBlockK = 2 * BLOCK_K;            // 64 halves per 128-byte swizzle row
  const int groups_per_comp = p.n_tiles / kPairs; // host: n_tiles % kPairs == 0
  const int units_per_m = 2 * groups_per_comp;    // (component, group of kPairs k tiles)
  const int n_units = p.m_tiles * units_per_m;
  const int kb_per_chain = p.quarter / kBlockK;
  const int num_kb = 2 * kb_per_chain;            // even-n chain, odd-n chain
  const int unit0 = (int)(blockIdx.x / (2 * kPairs)), unit_step = (int)(gridDim.x / (2 * kPairs));

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t leader_full0 = map_to_rank(bar_full(0), leader_rank);
      uint16_t same_half = 0;
#pragma unroll
      for (int j = 0; j < kPairs; ++j) same_half |= (uint16_t)(1u << (2 * j + (int)half_id));
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = unit0; unit < n_units; unit += unit_step) {
        const int m_tile = unit / units_per_m, rest = unit - m_tile * units_per_m;
        const int comp = rest / groups_per_comp, n_tile = (rest - comp * groups_per_comp) * kPairs + (int)pair;
        const int a_row = (int)(comp * p.m_rows) + m_tile * 256 + (int)half_id * 128 + (int)pair * kSliceRows;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int chain = kb >= kb_per_chain;
          const int kk = (kb - chain * kb_per_chain) * kBlockK;
          const int b_row = (2 * comp + chain) * p.n_k + n_tile * F_BLOCK_N + (int)half_id * 64;
          mbar_wait_cluster(bar_empty(stage), phase ^ 1u, nullptr, 1);
          if (leader) mbar_expect_tx(bar_full(stage), 2 * P_STAGE_BYTES);
          const uint32_t fb = leader_full0 + 8 * stage;
          tma_load_2d_pair_mc(&tm_a_hi, s_a(stage, 0) + pair * kSliceBytes, bar_full(stage), chain * p.quarter + kk, a_row,
                              same_half);
          tma_load_2d_pair_mc(&tm_a_lo, s_a(stage, 1) + pair * kSliceBytes, bar_full(stage), chain * p.quarter + kk, a_row,
                              same_half);
          tma_load_2d_pair(&tm_b_hi, s_b(stage, 0), fb, kk, b_row);
          tma_load_2d_pair(&tm_b_lo, s_b(stage, 1), fb, kk, b_row);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc(256, F_BLOCK_N, FMT_F16);
      constexpr uint16_t kAll = (uint16_t)((1u << (2 * kPairs)) - 1u);
      const uint16_t own_pair = (uint16_t)(3u << leader_rank);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int unit = unit0; unit < n_units; unit += unit_step) {
        mbar_wait(bar_tmem_empty(acc), acc_phase ^ 1u, nullptr, 2);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          const int chain = kb >= kb_per_chain;
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS + chain * F_BLOCK_N);
          const bool first_kb = (kb == 0) || (kb == kb_per_chain);
          mbar_wait(bar_full(stage), phase, nullptr, 3);
          tc_fence_after();
          const uint64_t da_hi = make_sw128_desc(s_a(stage, 0));
          const uint64_t da_lo = make_sw128_desc(s_a(stage, 1));
          const uint64_t db_hi = make_sw128_desc(s_b(stage, 0));
          const uint64_t db_lo = make_sw128_desc(s_b(stage, 1));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);             // one MMA consumes 32 bytes of the row
            umma_f16_pair(d_tmem, da_hi + adv, db_lo + adv, idesc, (first_kb && k == 0) ? 0u : 1u);
            umma_f16_pair(d_tmem, da_lo + adv, db_hi + adv, idesc, 1u);
            umma_f16_pair(d_tmem, da_hi + adv, db_hi + adv, idesc, 1u);
          }
          umma_commit_mask(bar_empty(stage), kAll);          // this pair is done with stage s: tell every producer
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_mask(bar_tmem_full(acc), own_pair);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    const int group = (warp - EPI_WARP0) >> 2;      // 0 .. 3
    const int stream = group >> 1, hhalf = group & 1;
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t leader_tmem_empty0 = map_to_rank(bar_tmem_empty(0), leader_rank);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = unit0; unit < n_units; unit += unit_step) {
      const int m_tile = unit / units_per_m, rest = unit - m_tile * units_per_m;
      const int comp = rest / groups_per_comp, n_tile = (rest - comp * groups_per_comp) * kPairs + (int)pair;
      const int64_t f = (int64_t)m_tile * 256 + (int64_t)half_id * 128 + row;    // flattened frame index
      const bool f_ok = f < p.m_rows;
      const int b = f_ok ? (int)(f / p.n_frames) : 0;
      const int t = f_ok ? (int)(f - (int64_t)b * p.n_frames) : 0;
      const float scale = f_ok ? __ldg(p.row_scale_inv + f) * p.basis_scale_inv : 1.f;
      float* col = p.mel_out + comp * p.plane_stride + (int64_t)b * p.n_mels * p.n_frames + t;
      const float4* tt = tab.e + stream * p.n_k + n_tile * F_BLOCK_N + hhalf * C_CHUNK;
      mbar_wait(bar_tmem_full(acc), acc_phase, nullptr, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_COLS + hhalf * C_CHUNK);
      mel2c_unit(p, taddr, tt, stream, col, f_ok, scale);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tmem_empty0 + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                             // peers may still be writing our smem / signalling our barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------- twice-folded kernel with the fold done IN the kernel (K1x)
// K1q re-reads 168 MB of materialised frame planes (every sample stored 4 times as fp16 hi/lo of e and o) that K0q
// has to write first.  Here the A operand is produced on the fly: eight CONVERTER warps per CTA read the raw padded
// PCM16 signal (parity planes of rvb_pad_parity_pcm16, 2 bytes per sample, L2-resident after the first touch), form
//     e[n] = p[n] + p[N-n]   or   o[n] = p[n] - p[N-n]          (this unit's component)
// exactly, split the value into fp16 hi + lo (exact: 17-bit integers) and store the two 128-row x 64-column tiles of
// a stage straight into the 128-byte-swizzled layout the MMA descriptors expect -- what the TMA did in K1q.  The
// even-n and the odd-n chain of a 128-sample block come from the same loads, so stages alternate between the chains:
// stage 2i = even chain, block i; stage 2i+1 = odd chain, block i.  The TMA warp only fetches the basis tiles.
//   full[s]   (leader)  1 arrive.expect_tx of the leader's producer (basis bytes of both CTAs) + one arrival per
//                       converter warp of BOTH CTAs (remote for rank 1), each after fence.proxy.async: the generic-
//                       proxy stores to shared memory are ordered before the async-proxy reads of tcgen05.mma
//   empty[s]  (per CTA) as in K1q (multicast commit); the converter warps wait on it like the producer does
// Sample -> float without a conversion instruction: the planes hold u = x + 32768; PRMT glues the 16 bits under the
// constant 0x4AC0: the float 1.5 * 2^22 + u / 2.  o / 2 = f(ux) - f(uy) exactly; e / 2 = f(ux) - f(~uy) - 1/2
// (~u = 65535 - u).  Column j of the even chain is n = 2 j + 2 (x = even-plane element 256 t + j + 1, one element past
// the 16-byte alignment: taken from the neighbouring lane), of the odd chain n = 2 j + 1; the partner p[N-n] is element
// 256 t + 1023 - j of the same plane for both.  The centre sample (n = N/2, even chain, last column) has no partner:
// e = 2 p[N/2] here, and the basis carries HALF the weight in that column (basis.fold2_operand(centre_doubled=True)).
constexpr int X_CONV_WARPS = 8;
constexpr int X_NUM_THREADS = P_NUM_THREADS + 32 * X_CONV_WARPS;      // 832
constexpr int X_CONV_WARP0 = P_NUM_THREADS / 32;                      // first converter warp

struct FusedParams {
  int n_frames;               // frames per segment (T)
  int64_t m_rows;             // n_seg * n_frames
  int n_k, quarter;           // basis rows per chain (N/4), contraction length per chain (N/4)
  int m_tiles, n_tiles;
  const uint16_t* sig;        // [2][n_seg][plane_len] parity planes, offset binary
  int64_t plane_len;
  int n_seg, hop2;            // hop / 2: plane elements between the starts of consecutive frames
  float scale;                // 2 * gain * basis_scale_inv: undoes the e/2 representation and the block scaling of the basis
  float* mel_out;             // [2][n_seg][n_mels][n_frames]
  int64_t plane_stride;
  int n_mels;
};

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Two columns (x elements in xr, low half first; partner elements in yr, HIGH half first) -> packed fp16 hi and lo.
__device__ __forceinline__ void fold_pair(uint32_t xr, uint32_t yr, float bias, uint32_t& hi, uint32_t& lo) {
  constexpr uint32_t kMagic = 0x4AC00000u;                 // 1.5 * 2^22: the low 16 mantissa bits weigh 1/2 .. 2^14
  const float x0 = __uint_as_float(__byte_perm(xr, kMagic, 0x7610));
  const float x1 = __uint_as_float(__byte_perm(xr, kMagic, 0x7632));
  const float y0 = __uint_as_float(__byte_perm(yr, kMagic, 0x7632));
  const float y1 = __uint_as_float(__byte_perm(yr, kMagic, 0x7610));
  const float v0 = (x0 - y0) - bias, v1 = (x1 - y1) - bias;           // exact: multiples of 1/2 below 2^16
  const __half2 h = __floats2half2_rn(v0, v1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);          // exact: |v - hi| <= 16
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <int kStages>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(X_NUM_THREADS, 1)
stft_gemm_fold2x_pair_kernel(const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                             const FusedParams p, const __grid_constant__ MelTable tab) {
  static_assert(kStages % 2 == 0, "stages are filled in (even chain, odd chain) pairs");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * P_STAGE_BYTES;
  auto s_a = [&](int s, int lo) { return smem_base + s * P_STAGE_BYTES + lo * P_A_BYTES; };
  auto s_b = [&](int s, int lo) { return smem_base + s * P_STAGE_BYTES + 2 * P_A_BYTES + lo * P_B_BYTES; };
  auto bar_full = [&](int s) { return bar_base + 8 * s; };
  auto bar_empty = [&](int s) { return bar_base + 8 * (kStages + s); };
  auto bar_tmem_full = [&](int a) { return bar_base + 8 * (2 * kStages + a); };
  auto bar_tmem_empty = [&](int a) { return bar_base + 8 * (2 * kStages + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8 * (2 * kStages + 4) + 16;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_b_hi);
    tma_prefetch_desc(&tm_b_lo);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full(s), 1 + 2 * X_CONV_WARPS);
      mbar_init(bar_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tmem_full(a), 1);
      mbar_init(bar_tmem_empty(a), 8 * P_EPI_GROUPS);
    }
    fence_barrier_init();
  }
  __syncthreads();                                     // barriers initialised before the allocator touches its slot
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr) : "memory");

  constexpr int kBlockK = 2 * BLOCK_K;            // 64 halves per 128-byte swizzle row
  const int units_per_m = 2 * p.n_tiles;          // (component, 128-k tile)
  const int n_units = p.m_tiles * units_per_m;
  const int n_blocks = p.quarter / kBlockK;       // 128-sample blocks per frame half = stage pairs per unit
  const int num_kb = 2 * n_blocks;
  const int unit0 = (int)(blockIdx.x >> 1), unit_step = (int)(gridDim.x >> 1);

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t leader_full0 = map_to_rank(bar_full(0), 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = unit0; unit < n_units; unit += unit_step) {
        const int rest = unit % units_per_m;
        const int comp = rest / p.n_tiles, n_tile = rest - comp * p.n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int chain = kb & 1;
          const int kk = (kb >> 1) * kBlockK;
          const int b_row = (2 * comp + chain) * p.n_k + n_tile * F_BLOCK_N + (int)rank * 64;
          mbar_wait(bar_empty(stage), phase ^ 1u, nullptr, 1);
          if (leader) mbar_expect_tx(bar_full(stage), 2 * 2 * P_B_BYTES);
          const uint32_t fb = leader_full0 + 8 * stage;
          tma_load_2d_pair(&tm_b_hi, s_b(stage, 0), fb, kk, b_row);
          tma_load_2d_pair(&tm_b_lo, s_b(stage, 1), fb, kk, b_row);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc(256, F_BLOCK_N, FMT_F16);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int unit = unit0; unit < n_units; unit += unit_step) {
        mbar_wait(bar_tmem_empty(acc), acc_phase ^ 1u, nullptr, 2);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          const int chain = kb & 1;
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS + chain * F_BLOCK_N);
          mbar_wait_cluster(bar_full(stage), phase, nullptr, 3);
          tc_fence_after();
          const uint64_t da_hi = make_sw128_desc(s_a(stage, 0));
          const uint64_t da_lo = make_sw128_desc(s_a(stage, 1));
          const uint64_t db_hi = make_sw128_desc(s_b(stage, 0));
          const uint64_t db_lo = make_sw128_desc(s_b(stage, 1));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);             // one MMA consumes 32 bytes of the row
            umma_f16_pair(d_tmem, da_hi + adv, db_lo + adv, idesc, (kb < 2 && k == 0) ? 0u : 1u);
            umma_f16_pair(d_tmem, da_lo + adv, db_hi + adv, idesc, 1u);
            umma_f16_pair(d_tmem, da_hi + adv, db_hi + adv, idesc, 1u);
          }
          umma_commit_pair(bar_empty(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_pair(bar_tmem_full(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp >= X_CONV_WARP0) {
    // ---- converter warps: raw samples -> folded, split, swizzled A tiles of a stage pair
    const int cw = warp - X_CONV_WARP0;
    const int q = lane >> 3, sub = lane & 7;              // quarter-warp = (row, chain); lane = 16-byte chunk of the row
    const int chain = q & 1;
    const int sh = chain ? 0 : 16;                        // even chain: x starts one element past the chunk
    const uint32_t leader_full0 = map_to_rank(bar_full(0), 0);
    const uint16_t* plane = p.sig + (int64_t)chain * p.n_seg * p.plane_len;
    int stage = 0;
    uint32_t phase = 0;
    for (int unit = unit0; unit < n_units; unit += unit_step) {
      const int m_tile = unit / units_per_m, rest = unit - m_tile * units_per_m;
      const int comp = rest / p.n_tiles;
      const uint32_t ymask = comp ? 0u : 0xffffffffu;     // e: partner complemented (f(ux) - f(~uy) = (x + y + 1) / 2)
      const float bias = comp ? 0.f : 0.5f;
      int32_t off[8];                                     // plane offset of the frame start of this lane's 8 rows
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = it * 16 + cw * 2 + (q >> 1);
        int64_t f = (int64_t)m_tile * 256 + (int64_t)rank * 128 + row;
        if (f >= p.m_rows) f = p.m_rows - 1;              // rows past the end: any valid frame (the epilogue drops them)
        const int b = (int)(f / p.n_frames);
        const int t = (int)(f - (int64_t)b * p.n_frames);
        off[it] = (int32_t)((int64_t)b * p.plane_len + (int64_t)t * p.hop2);
      }
      // The loads of a row are issued kPrefetch rows ahead of its conversion, across the block (stage pair) boundary:
      // they do not depend on the empty barrier, only the shared-memory stores do.  Little's law: 8 warps x 3 rows x
      // 1.1 KB in flight per SM (27 KB) cover ~600 cycles of L2 latency at the 26 bytes/clock the MMA pace asks for;
      // one row ahead (the first version) left the converters latency-bound at half that rate.
      constexpr int kPrefetch = 3;
      uint4 X[4], Y[4];
      uint32_t XN[4];
      const int n_rows_unit = 8 * n_blocks;
      auto issue = [&](int slot, int it, int blk) {
        const int xo = blk * kBlockK + sub * 8, yo = p.quarter * 2 - 8 - xo;     // N/2 - 8 - xo
        const uint16_t* base = plane + off[it];
        X[slot] = ldg_nc_v4(base + xo);
        Y[slot] = ldg_nc_v4(base + yo);
        XN[slot] = (sub == 7) ? __ldg(reinterpret_cast<const uint32_t*>(base + xo + 8)) : 0u;
      };
#pragma unroll
      for (int j = 0; j < kPrefetch; ++j) issue(j, j, 0);
      for (int blk = 0; blk < n_blocks; ++blk) {
        mbar_wait(bar_empty(stage), phase ^ 1u, nullptr, 5);
        mbar_wait(bar_empty(stage + 1), phase ^ 1u, nullptr, 5);
        const int st = stage + chain;                     // this quarter-warp's stage of the pair
        const uint32_t a_hi = s_a(st, 0), a_lo = s_a(st, 1);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int cur = it & 3;
          const uint4 Xc = X[cur], Yc = Y[cur];
          const uint32_t XNc = XN[cur];
          {
            constexpr int kAhead = kPrefetch;
            const int itn = (it + kAhead) & 7, blkn = blk + ((it + kAhead) >> 3);
            if (blk * 8 + it + kAhead < n_rows_unit) issue((it + kAhead) & 3, itn, blkn);
          }
          // element 8 of this chunk = element 0 of the next lane's chunk (same row for sub < 7)
          uint32_t nx = __shfl_down_sync(kFull, Xc.x, 1);
          if (sub == 7) nx = XNc;
          const uint32_t x0 = __funnelshift_r(Xc.x, Xc.y, sh), x1 = __funnelshift_r(Xc.y, Xc.z, sh);
          const uint32_t x2 = __funnelshift_r(Xc.z, Xc.w, sh), x3 = __funnelshift_r(Xc.w, nx, sh);
          uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
          fold_pair(x0, Yc.w ^ ymask, bias, h0, l0);     // columns 8 sub + 0, 1  <->  partner elements 7, 6
          fold_pair(x1, Yc.z ^ ymask, bias, h1, l1);
          fold_pair(x2, Yc.y ^ ymask, bias, h2, l2);
          fold_pair(x3, Yc.x ^ ymask, bias, h3, l3);
          const int row = it * 16 + cw * 2 + (q >> 1);
          const uint32_t dst = (uint32_t)(row * 128 + ((sub ^ (row & 7)) << 4));
          sts_v4(a_hi + dst, h0, h1, h2, h3);
          sts_v4(a_lo + dst, l0, l1, l2, l3);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(leader_full0 + 8 * stage);
          mbar_arrive_cluster(leader_full0 + 8 * (stage + 1));
        }
        stage += 2;
        if (stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    const int group = (warp - EPI_WARP0) >> 2;      // 0 .. 3
    const int stream = group >> 1, hhalf = group & 1;
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t leader_tmem_empty0 = map_to_rank(bar_tmem_empty(0), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    Fold2Params ep;                                  // the epilogue of K1q, fed with this kernel's geometry
    ep.n_frames = p.n_frames; ep.n_mels = p.n_mels;
    for (int unit = unit0; unit < n_units; unit += unit_step) {
      const int m_tile = unit / units_per_m, rest = unit - m_tile * units_per_m;
      const int comp = rest / p.n_tiles, n_tile = rest - comp * p.n_tiles;
      const int64_t f = (int64_t)m_tile * 256 + (int64_t)rank * 128 + row;       // flattened frame index
      const bool f_ok = f < p.m_rows;
      const int b = f_ok ? (int)(f / p.n_frames) : 0;
      const int t = f_ok ? (int)(f - (int64_t)b * p.n_frames) : 0;
      float* col = p.mel_out + comp * p.plane_stride + (int64_t)b * p.n_mels * p.n_frames + t;
      const float4* tt = tab.e + stream * p.n_k + n_tile * F_BLOCK_N + hhalf * C_CHUNK;
      mbar_wait(bar_tmem_full(acc), acc_phase, nullptr, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_COLS + hhalf * C_CHUNK);
      mel2c_unit(ep, taddr, tt, stream, col, f_ok, p.scale);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tmem_empty0 + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// Single bin from the folded planes (Nyquist bin of the STFT module): warp per frame, fp32 FMA.
// T = float (tf32 planes) or __half (fp16 planes, row-scaled: row_scale_inv undoes the scaling).
template <typename T>
__global__ void __launch_bounds__(256)
stft_bin_fold_kernel(const T* __restrict__ a_hi, const T* __restrict__ a_lo, const float* __restrict__ row_scale_inv,
                     int64_t m_rows, int n_frames, int half, const float* __restrict__ wc_row,
                     const float* __restrict__ ws_row, const float* __restrict__ p0, float w0, int bin, int epilogue,
                     float power, float* __restrict__ out0, int n_out_bins) {
  const int lane = threadIdx.x & 31;
  const int64_t f = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (f >= m_rows) return;
  const T* e_hi = a_hi + f * half;
  const T* e_lo = a_lo + f * half;
  const T* o_hi = a_hi + (m_rows + f) * half;
  const T* o_lo = a_lo + (m_rows + f) * half;
  float re = 0.f, im = 0.f;
  for (int c = lane; c < half; c += 32) {
    re = fmaf((float)e_hi[c] + (float)e_lo[c], __ldg(wc_row + c), re);
    im = fmaf((float)o_hi[c] + (float)o_lo[c], __ldg(ws_row + c), im);
  }
  re = warp_sum(re);
  im = warp_sum(im);
  if (lane == 0) {
    if (row_scale_inv) {
      const float sc = __ldg(row_scale_inv + f);
      re *= sc; im *= sc;
    }
    if (p0) re += w0 * __ldg(p0 + f);
    const int b = (int)(f / n_frames), t = (int)(f - (int64_t)b * n_frames);
    stft_store(epilogue, power, re, im, out0, b, bin, t, n_out_bins, n_frames);
  }
}

// ---------------------------------------------------------------- host side
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  return fn;
}

// 2-D row-major tensor [rows][cols] of fp32 (elem_bytes 4) or fp16 (2) -> tiled map with a
// (box_cols x box_rows) box, 128-byte swizzle.
static int make_map_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint32_t box_cols,
                       uint32_t box_rows, int elem_bytes = 4) {
  using Key = std::tuple<const void*, uint64_t, uint64_t, uint32_t, uint32_t, int>;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  const Key key{base, cols, rows, box_cols, box_rows, elem_bytes};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return RVB_OK;
    }
  }
  auto encode = get_encode_fn();
  if (!encode) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return RVB_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = encode(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                      const_cast<void*>(base), gdim, gstride, box, estride,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (cols=%llu rows=%llu box=%ux%u)", (int)r,
              (unsigned long long)cols, (unsigned long long)rows, box_cols, box_rows);
    return RVB_ERR_CUDA;
  }
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 256) cache.clear();
  cache[key] = *out;
  return RVB_OK;
}

static int num_sms() {
  static int n[kMaxDevices] = {};
  const int slot = device_slot();
  std::lock_guard<std::mutex> g(attr_mutex());
  if (n[slot] == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n[slot], cudaDevAttrMultiProcessorCount, dev);
    if (n[slot] <= 0) n[slot] = 148;
  }
  return n[slot];
}

}  // namespace rvb

using namespace rvb;

static int launch_stft_gemm(const CUtensorMap& tm_a_hi, const CUtensorMap& tm_a_lo, const CUtensorMap& tm_b_hi,
                            const CUtensorMap& tm_b_lo, const GemmParams& p, rvb_stream_t stream) {
  {
    static bool attr_set[kMaxDevices] = {};
    const int slot = device_slot();
    std::lock_guard<std::mutex> g(attr_mutex());
    if (!attr_set[slot]) {
      RVB_CUDA(cudaFuncSetAttribute(stft_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      attr_set[slot] = true;
    }
  }
  const int64_t n_units = (int64_t)p.m_tiles * p.n_tiles * p.k_split;
  const int grid = (int)(n_units < num_sms() ? n_units : num_sms());
  stft_gemm_kernel<<<grid, NUM_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo, p);
  count_launch();
  return check_launch("stft_gemm_kernel");
}

extern "C" int rvb_stft_gemm(const float* sig_hi, const float* sig_lo, int n_seg, int rows_per_seg, int hop,
                             int n_frames, const float* basis_hi, const float* basis_lo, int n_basis_rows, int n_fft,
                             int epilogue, float power, float* out0, int n_out_bins, rvb_stream_t stream) {
  RVB_REQUIRE(sig_hi && sig_lo && basis_hi && basis_lo && out0, "rvb_stft_gemm: null pointer");
  RVB_REQUIRE(n_seg > 0 && n_frames > 0 && rows_per_seg > 0, "rvb_stft_gemm: bad shape");
  RVB_REQUIRE(n_fft % BLOCK_K == 0 && n_fft >= BLOCK_K, "rvb_stft_gemm: n_fft %d must be a multiple of %d", n_fft,
              BLOCK_K);
  RVB_REQUIRE(hop % BLOCK_K == 0 && hop >= BLOCK_K, "rvb_stft_gemm: hop %d must be a multiple of %d", hop, BLOCK_K);
  RVB_REQUIRE(n_basis_rows % BLOCK_N == 0 && n_basis_rows > 0, "rvb_stft_gemm: n_basis_rows %d must be a multiple of %d",
              n_basis_rows, BLOCK_N);
  RVB_REQUIRE(epilogue_ok(epilogue), "rvb_stft_gemm: bad epilogue %d", epilogue);
  RVB_REQUIRE((int64_t)(n_frames - 1) * hop + n_fft <= (int64_t)rows_per_seg * hop,
              "rvb_stft_gemm: %d frames of %d samples do not fit %d rows of %d", n_frames, n_fft, rows_per_seg, hop);
  for (const void* ptr : {(const void*)sig_hi, (const void*)sig_lo, (const void*)basis_hi, (const void*)basis_lo})
    RVB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 127u) == 0, "rvb_stft_gemm: operands must be 128-byte aligned");
  RVB_REQUIRE(n_out_bins > 0, "rvb_stft_gemm: n_out_bins must be positive");

  CUtensorMap tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo;
  const uint64_t a_rows = (uint64_t)n_seg * rows_per_seg;
  int rc;
  if ((rc = make_map_2d(&tm_a_hi, sig_hi, hop, a_rows, BLOCK_K, BLOCK_M)) != RVB_OK) return rc;
  if ((rc = make_map_2d(&tm_a_lo, sig_lo, hop, a_rows, BLOCK_K, BLOCK_M)) != RVB_OK) return rc;
  if ((rc = make_map_2d(&tm_b_hi, basis_hi, n_fft, n_basis_rows, BLOCK_K, BLOCK_N)) != RVB_OK) return rc;
  if ((rc = make_map_2d(&tm_b_lo, basis_lo, n_fft, n_basis_rows, BLOCK_K, BLOCK_N)) != RVB_OK) return rc;

  GemmParams p;
  p.n_seg = n_seg; p.rows_per_seg = rows_per_seg; p.hop = hop; p.n_frames = n_frames; p.n_fft = n_fft;
  p.tiles_per_seg = (n_frames + BLOCK_M - 1) / BLOCK_M;
  p.m_tiles = p.tiles_per_seg * n_seg;
  p.n_tiles = n_basis_rows / BLOCK_N;
  p.epilogue = epilogue; p.n_out_bins = n_out_bins;
  p.n_store_bins = n_out_bins < n_basis_rows / 2 ? n_out_bins : n_basis_rows / 2;
  p.power = power; p.out0 = out0; p.dbg_status = nullptr; p.ldc = 0; p.split_stride = 0; p.k_split = 1;
  return launch_stft_gemm(tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo, p, stream);
}

// C[M][N] = A[M][K] . B[N][K]^T in 3xTF32 (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM): the dense projections of
// the caller's sequence model (nn.Linear of MutliHeadAttention1D, model/self_attention_VAT.py:54-56, 70-71), which
// PyTorch runs as SIMT SGEMMs (TF32 is off by default for matmul).  Operands: tf32 hi / lo planes [rows][k_pad] from
// rvb_split_tf32, k_pad a multiple of 32, zero beyond K.  Rows past M / N are zero-filled by the TMA and never stored.
// k_split > 1: the contraction is cut into k_split slices of whole 32-element blocks and slice s writes its partial product
// to c + s * split_stride -- for products with few output tiles and a long contraction (dW = dY^T X: 916 x 229 outputs,
// 20 480 terms); the caller adds the partials in a fixed order (deterministic, unlike atomics).
extern "C" int rvb_gemm_nt_tf32x3(const float* a_hi, const float* a_lo, int64_t m, const float* b_hi, const float* b_lo,
                                  int n, int k_pad, float* c, int64_t ldc, int k_split, int64_t split_stride,
                                  rvb_stream_t stream) {
  RVB_REQUIRE(a_hi && a_lo && b_hi && b_lo && c, "rvb_gemm_nt_tf32x3: null pointer");
  RVB_REQUIRE(m > 0 && n > 0 && k_pad >= BLOCK_K && k_pad % BLOCK_K == 0 && ldc >= n,
              "rvb_gemm_nt_tf32x3: bad shape (m=%lld n=%d k_pad=%d ldc=%lld)", (long long)m, n, k_pad, (long long)ldc);
  RVB_REQUIRE(m < (1ll << 31) - BLOCK_M, "rvb_gemm_nt_tf32x3: too many rows");
  RVB_REQUIRE(k_split >= 1 && k_split <= k_pad / BLOCK_K && (k_split == 1 || split_stride >= m * ldc),
              "rvb_gemm_nt_tf32x3: bad split (k_split=%d of %d blocks, stride %lld)", k_split, k_pad / BLOCK_K,
              (long long)split_stride);
  {
    // no empty slice: with kb_per = ceil(blocks / k_split), slice k_split - 1 must still start inside the contraction
    const int blocks = k_pad / BLOCK_K, kb_per = (blocks + k_split - 1) / k_split;
    RVB_REQUIRE((k_split - 1) * kb_per < blocks, "rvb_gemm_nt_tf32x3: k_split %d leaves an empty slice of %d blocks", k_split, blocks);
  }
  for (const void* ptr : {(const void*)a_hi, (const void*)a_lo, (const void*)b_hi, (const void*)b_lo})
    RVB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 127u) == 0, "rvb_gemm_nt_tf32x3: operands must be 128-byte aligned");
  CUtensorMap tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo;
  int rc;
  if ((rc = make_map_2d(&tm_a_hi, a_hi, k_pad, (uint64_t)m, BLOCK_K, BLOCK_M)) != RVB_OK) return rc;
  if ((rc = make_map_2d(&tm_a_lo, a_lo, k_pad, (uint64_t)m, BLOCK_K, BLOCK_M)) != RVB_OK) return rc;
  if ((rc = make_map_2d(&tm_b_hi, b_hi, k_pad, (uint64_t)n, BLOCK_K, BLOCK_N)) != RVB_OK) return rc;
  if ((rc = make_map_2d(&tm_b_lo, b_lo, k_pad, (uint64_t)n, BLOCK_K, BLOCK_N)) != RVB_OK) return rc;
  GemmParams p;
  p.n_seg = 1; p.rows_per_seg = (int)m; p.hop = k_pad; p.n_frames = (int)m; p.n_fft = k_pad;
  p.tiles_per_seg = (int)((m + BLOCK_M - 1) / BLOCK_M);
  p.m_tiles = p.tiles_per_seg;
  p.n_tiles = (n + BLOCK_N - 1) / BLOCK_N;
  p.epilogue = kEpiRawGemm; p.n_out_bins = n; p.n_store_bins = n;
  p.power = 0.f; p.out0 = c; p.dbg_status = nullptr; p.ldc = ldc; p.k_split = k_split; p.split_stride = split_stride;
  return launch_stft_gemm(tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo, p, stream);
}

struct MelArgs {
  const float* tab;   // [n_bins_pad][4]
  float* out;         // [n_seg][n_mels][n_frames]
  int n_mels;
};

template <bool kF16>
static int launch_folded(const char* who, const void* a_hi, const void* a_lo, const float* row_scale_inv, int n_seg,
                         int n_frames, int n_fft, const void* basis_hi, const void* basis_lo, float basis_scale_inv,
                         int n_bins_pad, const float* p0, float w0, int epilogue, float power, float* out0,
                         int n_out_bins, rvb_stream_t stream, const MelArgs* mel = nullptr) {
  constexpr int kBlockK = kF16 ? 2 * BLOCK_K : BLOCK_K;
  constexpr int kElem = kF16 ? 2 : 4;
  RVB_REQUIRE(a_hi && a_lo && basis_hi && basis_lo && (out0 || mel), "%s: null pointer", who);
  RVB_REQUIRE(!kF16 || row_scale_inv, "%s: null row_scale_inv", who);
  if (mel) {
    RVB_REQUIRE(mel->tab && mel->out && mel->n_mels > 0, "%s: bad Mel arguments", who);
    RVB_REQUIRE(n_bins_pad <= kMelTableBins, "%s: the Mel table is a %d-bin kernel parameter, got n_bins_pad %d", who,
                kMelTableBins, n_bins_pad);
    RVB_REQUIRE(epilogue == RVB_EPI_POWER || epilogue == RVB_EPI_MAGNITUDE || epilogue == RVB_EPI_POWER_P,
                "%s: the Mel projection takes the power / magnitude / power_p spectrum, got %d", who, epilogue);
    RVB_CUDA(cudaMemsetAsync(mel->out, 0, sizeof(float) * (size_t)n_seg * mel->n_mels * n_frames, (cudaStream_t)stream));
  }
  RVB_REQUIRE(n_seg > 0 && n_frames > 0, "%s: bad shape", who);
  RVB_REQUIRE(n_fft % (2 * kBlockK) == 0 && n_fft >= 2 * kBlockK, "%s: n_fft %d must be a multiple of %d", who, n_fft,
              2 * kBlockK);
  RVB_REQUIRE(n_bins_pad % F_BLOCK_N == 0 && n_bins_pad > 0, "%s: n_bins_pad %d must be a multiple of %d", who,
              n_bins_pad, F_BLOCK_N);
  RVB_REQUIRE(epilogue_ok(epilogue), "%s: bad epilogue %d", who, epilogue);
  RVB_REQUIRE(n_out_bins > 0, "%s: n_out_bins must be positive", who);
  RVB_REQUIRE(w0 == 0.f || p0 != nullptr, "%s: w0 != 0 needs p0", who);
  for (const void* ptr : {a_hi, a_lo, basis_hi, basis_lo})
    RVB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 127u) == 0, "%s: operands must be 128-byte aligned", who);
  const int half = n_fft / 2;
  const int64_t m_rows = (int64_t)n_seg * n_frames;
  RVB_REQUIRE(2 * m_rows < (1ll << 31), "%s: too many frames", who);

  CUtensorMap tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo;
  int rc;
  if ((rc = make_map_2d(&tm_a_hi, a_hi, half, 2 * m_rows, kBlockK, BLOCK_M, kElem)) != RVB_OK) return rc;
  if ((rc = make_map_2d(&tm_a_lo, a_lo, half, 2 * m_rows, kBlockK, BLOCK_M, kElem)) != RVB_OK) return rc;
  if ((rc = make_map_2d(&tm_b_hi, basis_hi, half, 2 * (uint64_t)n_bins_pad, kBlockK, F_BLOCK_N, kElem)) != RVB_OK) return rc;
  if ((rc = make_map_2d(&tm_b_lo, basis_lo, half, 2 * (uint64_t)n_bins_pad, kBlockK, F_BLOCK_N, kElem)) != RVB_OK) return rc;

  FoldParams p;
  p.n_frames = n_frames; p.m_rows = m_rows; p.n_bins_pad = n_bins_pad; p.half = half;
  p.m_tiles = (int)((m_rows + BLOCK_M - 1) / BLOCK_M);
  p.n_tiles = n_bins_pad / F_BLOCK_N;
  p.epilogue = epilogue; p.n_out_bins = n_out_bins;
  p.n_store_bins = n_out_bins < n_bins_pad ? n_out_bins : n_bins_pad;
  p.power = power; p.w0 = w0; p.p0 = (w0 != 0.f) ? p0 : nullptr; p.out0 = out0;
  p.row_scale_inv = row_scale_inv; p.basis_scale_inv = basis_scale_inv;
  p.mel_out = mel ? mel->out : nullptr;
  p.n_mels = mel ? mel->n_mels : 0;
  static const bool one_pass = getenv("RVB_FOLD_ONE_PASS") != nullptr;    // A/B switch for measurements
  p.corr_first = one_pass ? 0 : 1;

  if constexpr (kF16) {
    static const bool one_cta = getenv("RVB_GEMM_1CTA") != nullptr;       // A/B switch for measurements
    if (!one_cta) {
      // CTA-pair kernel: 256-frame tiles; the A map keeps its 128-row box, the B box shrinks to this CTA's 64 rows
      if ((rc = make_map_2d(&tm_b_hi, basis_hi, half, 2 * (uint64_t)n_bins_pad, kBlockK, 64, kElem)) != RVB_OK) return rc;
      if ((rc = make_map_2d(&tm_b_lo, basis_lo, half, 2 * (uint64_t)n_bins_pad, kBlockK, 64, kElem)) != RVB_OK) return rc;
      p.m_tiles = (int)((m_rows + 255) / 256);
      static int max_clusters_dev[kMaxDevices] = {};
      const int sms = num_sms(), slot = device_slot();
      int max_clusters;
      {
        std::lock_guard<std::mutex> g(attr_mutex());
        if (max_clusters_dev[slot] == 0) {
          RVB_CUDA(cudaFuncSetAttribute(stft_gemm_fold_pair_kernel<NoTable>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES));
          RVB_CUDA(cudaFuncSetAttribute(stft_gemm_fold_pair_kernel<MelTable>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES));
          cudaLaunchConfig_t qc = {};
          qc.gridDim = dim3(sms & ~1u); qc.blockDim = dim3(P_NUM_THREADS); qc.dynamicSmemBytes = P_SMEM_BYTES;
          int nc = 0;
          RVB_CUDA(cudaOccupancyMaxActiveClusters(&nc, stft_gemm_fold_pair_kernel<MelTable>, &qc));
          RVB_REQUIRE(nc > 0, "%s: no CTA pair fits on this device", who);
          max_clusters_dev[slot] = nc < sms / 2 ? nc : sms / 2;
        }
        max_clusters = max_clusters_dev[slot];
      }
      const int64_t n_units = (int64_t)p.m_tiles * p.n_tiles;
      const int n_clusters = (int)(n_units < max_clusters ? n_units : max_clusters);
      if (mel) {
        // the table is read on the HOST here and travels in the launch's parameter buffer (a CUDA graph bakes it in)
        static thread_local MelTable tab;
        std::memset(&tab, 0, sizeof(tab));
        std::memcpy(tab.e, mel->tab, sizeof(float4) * (size_t)n_bins_pad);
        stft_gemm_fold_pair_kernel<MelTable><<<2 * n_clusters, P_NUM_THREADS, P_SMEM_BYTES, (cudaStream_t)stream>>>(
            tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo, p, tab);
      } else {
        stft_gemm_fold_pair_kernel<NoTable><<<2 * n_clusters, P_NUM_THREADS, P_SMEM_BYTES, (cudaStream_t)stream>>>(
            tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo, p, NoTable{0});
      }
      count_launch();
      return check_launch("stft_gemm_fold_pair_kernel");
    }
  }
  RVB_REQUIRE(!mel, "%s: the Mel epilogue lives in the CTA-pair kernel (unset RVB_GEMM_1CTA)", who);
  {
    static bool attr_set[kMaxDevices] = {};
    const int slot = device_slot();
    std::lock_guard<std::mutex> g(attr_mutex());
    if (!attr_set[slot]) {
      RVB_CUDA(cudaFuncSetAttribute(stft_gemm_fold_kernel<kF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_BYTES));
      attr_set[slot] = true;
    }
  }
  const int64_t n_units = (int64_t)p.m_tiles * p.n_tiles;
  const int grid = (int)(n_units < num_sms() ? n_units : num_sms());
  stft_gemm_fold_kernel<kF16><<<grid, NUM_THREADS, F_SMEM_BYTES, (cudaStream_t)stream>>>(tm_a_hi, tm_a_lo, tm_b_hi,
                                                                                            tm_b_lo, p);
  count_launch();
  return check_launch(kF16 ? "stft_gemm_fold_kernel<f16>" : "stft_gemm_fold_kernel<tf32>");
}

extern "C" int rvb_stft_gemm_folded(const float* a_hi, const float* a_lo, int n_seg, int n_frames, int n_fft,
                                    const float* basis_hi, const float* basis_lo, int n_bins_pad, const float* p0,
                                    float w0, int epilogue, float power, float* out0, int n_out_bins,
                                    rvb_stream_t stream) {
  return launch_folded<false>("rvb_stft_gemm_folded", a_hi, a_lo, nullptr, n_seg, n_frames, n_fft, basis_hi, basis_lo,
                              1.f, n_bins_pad, p0, w0, epilogue, power, out0, n_out_bins, stream);
}

extern "C" int rvb_stft_gemm_folded_f16(const void* a_hi, const void* a_lo, const float* row_scale_inv, int n_seg,
                                        int n_frames, int n_fft, const void* basis_hi, const void* basis_lo,
                                        float basis_scale_inv, int n_bins_pad, const float* p0, float w0, int epilogue,
                                        float power, float* out0, int n_out_bins, rvb_stream_t stream) {
  return launch_folded<true>("rvb_stft_gemm_folded_f16", a_hi, a_lo, row_scale_inv, n_seg, n_frames, n_fft, basis_hi,
                             basis_lo, basis_scale_inv, n_bins_pad, p0, w0, epilogue, power, out0, n_out_bins, stream);
}

extern "C" int rvb_stft_mel_folded_f16(const void* a_hi, const void* a_lo, const float* row_scale_inv, int n_seg,
                                       int n_frames, int n_fft, const void* basis_hi, const void* basis_lo,
                                       float basis_scale_inv, int n_bins_pad, const float* p0, float w0, int spectrum,
                                       float power, const float* mel_tab, int n_mels, float* mel_out,
                                       rvb_stream_t stream) {
  const MelArgs mel{mel_tab, mel_out, n_mels};
  return launch_folded<true>("rvb_stft_mel_folded_f16", a_hi, a_lo, row_scale_inv, n_seg, n_frames, n_fft, basis_hi,
                             basis_lo, basis_scale_inv, n_bins_pad, p0, w0, spectrum, power, nullptr, n_bins_pad, stream,
                             &mel);
}

extern "C" int rvb_stft_mel_folded2_f16(const void* a_hi, const void* a_lo, const float* row_scale_inv, int n_seg,
                                        int n_frames, int n_fft, const void* basis_hi, const void* basis_lo,
                                        float basis_scale_inv, const float* mel_tab, int n_mels, float* mel_out,
                                        rvb_stream_t stream) {
  const char* who = "rvb_stft_mel_folded2_f16";
  RVB_REQUIRE(a_hi && a_lo && row_scale_inv && basis_hi && basis_lo && mel_tab && mel_out, "%s: null pointer", who);
  RVB_REQUIRE(n_seg > 0 && n_frames > 0 && n_mels > 0, "%s: bad shape", who);
  RVB_REQUIRE(n_fft >= 512 && n_fft % 512 == 0 && 2 * (n_fft / 4) <= kMelTableBins,
              "%s: n_fft %d must be a multiple of 512 and at most %d", who, n_fft, 2 * kMelTableBins);
  for (const void* ptr : {a_hi, a_lo, basis_hi, basis_lo})
    RVB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 127u) == 0, "%s: operands must be 128-byte aligned", who);
  const int half = n_fft / 2, quarter = n_fft / 4;
  const int64_t m_rows = (int64_t)n_seg * n_frames;
  RVB_REQUIRE(2 * m_rows < (1ll << 31), "%s: too many frames", who);
  const int64_t plane = (int64_t)n_seg * n_mels * n_frames;
  RVB_CUDA(cudaMemsetAsync(mel_out, 0, sizeof(float) * 2 * (size_t)plane, (cudaStream_t)stream));
  // A/B switch for measurements: RVB_FOLD2_N64=1 runs the four-chain kernel (both components per unit, L2-bound)
  static const bool n64 = [] { const char* e = getenv("RVB_FOLD2_N64"); return e && *e && *e != '0'; }();

  CUtensorMap tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo;
  int rc;
  const uint32_t b_box_rows = n64 ? 32 : 64;
  if ((rc = make_map_2d(&tm_a_hi, a_hi, half, 2 * m_rows, 64, 128, 2)) != RVB_OK) return rc;
  if ((rc = make_map_2d(&tm_a_lo, a_lo, half, 2 * m_rows, 64, 128, 2)) != RVB_OK) return rc;
  if ((rc = make_map_2d(&tm_b_hi, basis_hi, quarter, 4 * (uint64_t)quarter, 64, b_box_rows, 2)) != RVB_OK) return rc;
  if ((rc = make_map_2d(&tm_b_lo, basis_lo, quarter, 4 * (uint64_t)quarter, 64, b_box_rows, 2)) != RVB_OK) return rc;

  Fold2Params p;
  p.n_frames = n_frames; p.m_rows = m_rows; p.n_k = quarter; p.quarter = quarter;
  p.m_tiles = (int)((m_rows + 255) / 256);
  p.n_tiles = quarter / (n64 ? Q_BLOCK_N : F_BLOCK_N);
  p.row_scale_inv = row_scale_inv; p.basis_scale_inv = basis_scale_inv;
  p.mel_out = mel_out; p.plane_stride = plane; p.n_mels = n_mels;
  p.exp = 0;
#ifdef RVB_ABLATION
  { const char* e = getenv("RVB_EXP"); p.exp = e ? atoi(e) : 0; }
#endif

  // RVB_FOLD2_STAGES=3: a 3-stage operand ring (144 KB of shared memory instead of 192 KB leaves room for blocks of the
  // HBM kernels of other streams beside a resident CTA)
  static const bool three = [] { const char* e = getenv("RVB_FOLD2_STAGES"); return e && atoi(e) == 3; }();
  static int max_clusters_dev[kMaxDevices][2] = {};
  const int smem = n64 ? Q_SMEM_BYTES : (three ? 3 : P_STAGES) * P_STAGE_BYTES + BAR_BYTES + 1024;
  auto kernel = n64 ? stft_gemm_fold2_pair_kernel
                    : (three ? stft_gemm_fold2c_pair_kernel<3> : stft_gemm_fold2c_pair_kernel<P_STAGES>);
  const int sms = num_sms(), slot = device_slot();
  int max_clusters;
  {
    std::lock_guard<std::mutex> g(attr_mutex());
    if (max_clusters_dev[slot][n64] == 0) {
      RVB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      cudaLaunchConfig_t qc = {};
      qc.gridDim = dim3(sms & ~1u); qc.blockDim = dim3(P_NUM_THREADS); qc.dynamicSmemBytes = smem;
      int nc = 0;
      RVB_CUDA(cudaOccupancyMaxActiveClusters(&nc, kernel, &qc));
      RVB_REQUIRE(nc > 0, "%s: no CTA pair fits on this device", who);
      max_clusters_dev[slot][n64] = nc < sms / 2 ? nc : sms / 2;
    }
    max_clusters = max_clusters_dev[slot][n64];
  }
  // RVB_FOLD2_MC=<2|4>: K1qm, <n> CTA pairs per cluster share the frame tiles through TMA multicast
  const int mc_pairs = [] { const char* e = getenv("RVB_FOLD2_MC"); return e ? atoi(e) : 0; }();   // read per launch (tests)
  if (!n64 && !three && (mc_pairs == 2 || mc_pairs == 4) && p.n_tiles % mc_pairs == 0) {
    const uint32_t slice_rows = 128 / mc_pairs;
    if ((rc = make_map_2d(&tm_a_hi, a_hi, half, 2 * m_rows, 64, slice_rows, 2)) != RVB_OK) return rc;
    if ((rc = make_map_2d(&tm_a_lo, a_lo, half, 2 * m_rows, 64, slice_rows, 2)) != RVB_OK) return rc;
    auto mc_kernel = mc_pairs == 4 ? stft_gemm_fold2c_mc_kernel<P_STAGES, 4> : stft_gemm_fold2c_mc_kernel<P_STAGES, 2>;
    static int mc_clusters_dev[kMaxDevices][2] = {};
    const int slot2 = device_slot();
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2 * mc_pairs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.blockDim = dim3(P_NUM_THREADS); cfg.dynamicSmemBytes = P_SMEM_BYTES; cfg.stream = (cudaStream_t)stream;
    int mc_clusters;
    {
      std::lock_guard<std::mutex> g(attr_mutex());
      if (mc_clusters_dev[slot2][mc_pairs == 4] == 0) {
        RVB_CUDA(cudaFuncSetAttribute(mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES));
        cfg.gridDim = dim3((sms / (2 * mc_pairs)) * 2 * mc_pairs);   // (num_sms() takes attr_mutex itself)
        int nc = 0;
        RVB_CUDA(cudaOccupancyMaxActiveClusters(&nc, mc_kernel, &cfg));
        RVB_REQUIRE(nc > 0, "%s: no cluster of %d CTAs fits on this device", who, 2 * mc_pairs);
        mc_clusters_dev[slot2][mc_pairs == 4] = nc;
      }
      mc_clusters = mc_clusters_dev[slot2][mc_pairs == 4];
    }
    static const int mc_cap = [] { const char* e = getenv("RVB_FOLD2_CLUSTERS"); return e ? atoi(e) : 0; }();
    if (mc_cap > 0 && mc_cap < mc_clusters) mc_clusters = mc_cap;
    const int64_t g_units = (int64_t)p.m_tiles * 2 * (p.n_tiles / mc_pairs);
    const int n_cl = (int)(g_units < mc_clusters ? g_units : mc_clusters);
    static thread_local MelTable mtab;
    std::memset(&mtab, 0, sizeof(mtab));
    std::memcpy(mtab.e, mel_tab, sizeof(float4) * (size_t)(2 * quarter));
    cfg.gridDim = dim3(n_cl * 2 * mc_pairs);
    RVB_CUDA(cudaLaunchKernelEx(&cfg, mc_kernel, tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo, p, mtab));
    count_launch();
    return check_launch("stft_gemm_fold2c_mc_kernel");
  }
  const int64_t n_units = (int64_t)p.m_tiles * p.n_tiles * (n64 ? 1 : 2);
  // RVB_FOLD2_CLUSTERS=<n>: fewer CTA pairs than fit (A/B: the kernel is bound by L2 -> SM operand traffic, not by the
  // number of tensor pipes, so SMs left free go to the HBM kernels of other streams)
  static const int cap = [] { const char* e = getenv("RVB_FOLD2_CLUSTERS"); return e ? atoi(e) : 0; }();
  if (cap > 0 && cap < max_clusters) max_clusters = cap;
  const int n_clusters = (int)(n_units < max_clusters ? n_units : max_clusters);
  static thread_local MelTable tab;
  std::memset(&tab, 0, sizeof(tab));
  std::memcpy(tab.e, mel_tab, sizeof(float4) * (size_t)(2 * quarter));
  // (Tried: highest launch priority for this kernel, so that freed SMs go to its CTAs before another stream's HBM
  // blocks -- 0.216 instead of 0.209 ms per step on three streams; not kept.)
  kernel<<<2 * n_clusters, P_NUM_THREADS, smem, (cudaStream_t)stream>>>(tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo, p, tab);
  count_launch();
  return check_launch(n64 ? "stft_gemm_fold2_pair_kernel" : "stft_gemm_fold2c_pair_kernel");
}

extern "C" int rvb_stft_mel_fused_pcm16(const uint16_t* planes, int64_t plane_len, int n_seg, int n_frames, int n_fft,
                                        int hop, float gain, const void* basis_hi, const void* basis_lo,
                                        float basis_scale_inv, const float* mel_tab, int n_mels, float* mel_out,
                                        rvb_stream_t stream) {
  const char* who = "rvb_stft_mel_fused_pcm16";
  RVB_REQUIRE(planes && basis_hi && basis_lo && mel_tab && mel_out, "%s: null pointer", who);
  RVB_REQUIRE(n_seg > 0 && n_frames > 0 && n_mels > 0, "%s: bad shape", who);
  RVB_REQUIRE(n_fft >= 512 && n_fft % 512 == 0 && 2 * (n_fft / 4) <= kMelTableBins,
              "%s: n_fft %d must be a multiple of 512 and at most %d", who, n_fft, 2 * kMelTableBins);
  RVB_REQUIRE(hop > 0 && hop % 16 == 0, "%s: hop %d must be a multiple of 16 (16-byte loads from the parity planes)", who, hop);
  RVB_REQUIRE(plane_len % 8 == 0 && plane_len >= (int64_t)(hop / 2) * (n_frames - 1) + n_fft / 2 + 8,
              "%s: plane_len %lld too short for %d frames (rvb_parity_plane_len)", who, (long long)plane_len, n_frames);
  RVB_REQUIRE(2 * (int64_t)n_seg * plane_len < (1ll << 31), "%s: signal too long for 32-bit plane offsets", who);
  for (const void* ptr : {(const void*)planes, basis_hi, basis_lo})
    RVB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 127u) == 0, "%s: operands must be 128-byte aligned", who);
  const int quarter = n_fft / 4;
  const int64_t m_rows = (int64_t)n_seg * n_frames;
  RVB_REQUIRE(2 * m_rows < (1ll << 31), "%s: too many frames", who);
  const int64_t plane = (int64_t)n_seg * n_mels * n_frames;
  RVB_CUDA(cudaMemsetAsync(mel_out, 0, sizeof(float) * 2 * (size_t)plane, (cudaStream_t)stream));

  CUtensorMap tm_b_hi, tm_b_lo;
  int rc;
  if ((rc = make_map_2d(&tm_b_hi, basis_hi, quarter, 4 * (uint64_t)quarter, 64, 64, 2)) != RVB_OK) return rc;
  if ((rc = make_map_2d(&tm_b_lo, basis_lo, quarter, 4 * (uint64_t)quarter, 64, 64, 2)) != RVB_OK) return rc;

  FusedParams p;
  p.n_frames = n_frames; p.m_rows = m_rows; p.n_k = quarter; p.quarter = quarter;
  p.m_tiles = (int)((m_rows + 255) / 256);
  p.n_tiles = quarter / F_BLOCK_N;
  p.sig = planes; p.plane_len = plane_len; p.n_seg = n_seg; p.hop2 = hop / 2;
  p.scale = 2.f * gain * basis_scale_inv;              // A holds e / 2 (o / 2) in units of one PCM step
  p.mel_out = mel_out; p.plane_stride = plane; p.n_mels = n_mels;

  constexpr int smem = P_STAGES * P_STAGE_BYTES + BAR_BYTES + 1024;
  auto kernel = stft_gemm_fold2x_pair_kernel<P_STAGES>;
  static int max_clusters_dev[kMaxDevices] = {};
  const int sms = num_sms(), slot = device_slot();
  int max_clusters;
  {
    std::lock_guard<std::mutex> g(attr_mutex());
    if (max_clusters_dev[slot] == 0) {
      RVB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      cudaLaunchConfig_t qc = {};
      qc.gridDim = dim3(sms & ~1u); qc.blockDim = dim3(X_NUM_THREADS); qc.dynamicSmemBytes = smem;
      int nc = 0;
      RVB_CUDA(cudaOccupancyMaxActiveClusters(&nc, kernel, &qc));
      RVB_REQUIRE(nc > 0, "%s: no CTA pair fits on this device", who);
      max_clusters_dev[slot] = nc < sms / 2 ? nc : sms / 2;
    }
    max_clusters = max_clusters_dev[slot];
  }
  const int64_t n_units = (int64_t)p.m_tiles * p.n_tiles * 2;
  const int n_clusters = (int)(n_units < max_clusters ? n_units : max_clusters);
  static thread_local MelTable tab;
  std::memset(&tab, 0, sizeof(tab));
  std::memcpy(tab.e, mel_tab, sizeof(float4) * (size_t)(2 * quarter));
  kernel<<<2 * n_clusters, X_NUM_THREADS, smem, (cudaStream_t)stream>>>(tm_b_hi, tm_b_lo, p, tab);
  count_launch();
  return check_launch("stft_gemm_fold2x_pair_kernel");
}

template <typename T>
static int launch_bin_folded(const char* who, const T* a_hi, const T* a_lo, const float* row_scale_inv, int n_seg,
                             int n_frames, int n_fft, const float* wc_row, const float* ws_row, const float* p0,
                             float w0, int bin, int epilogue, float power, float* out0, int n_out_bins,
                             rvb_stream_t stream) {
  RVB_REQUIRE(a_hi && a_lo && wc_row && ws_row && out0, "%s: null pointer", who);
  RVB_REQUIRE(n_seg > 0 && n_frames > 0 && bin >= 0 && bin < n_out_bins, "%s: bad shape", who);
  RVB_REQUIRE(epilogue_ok(epilogue), "%s: bad epilogue %d", who, epilogue);
  RVB_REQUIRE(w0 == 0.f || p0 != nullptr, "%s: w0 != 0 needs p0", who);
  const int64_t m_rows = (int64_t)n_seg * n_frames;
  stft_bin_fold_kernel<T><<<(unsigned)((m_rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      a_hi, a_lo, row_scale_inv, m_rows, n_frames, n_fft / 2, wc_row, ws_row, (w0 != 0.f) ? p0 : nullptr, w0, bin,
      epilogue, power, out0, n_out_bins);
  count_launch();
  return check_launch("stft_bin_fold_kernel");
}

extern "C" int rvb_stft_bin_folded(const float* a_hi, const float* a_lo, int n_seg, int n_frames, int n_fft,
                                   const float* wc_row, const float* ws_row, const float* p0, float w0, int bin,
                                   int epilogue, float power, float* out0, int n_out_bins, rvb_stream_t stream) {
  return launch_bin_folded<float>("rvb_stft_bin_folded", a_hi, a_lo, nullptr, n_seg, n_frames, n_fft, wc_row, ws_row,
                                  p0, w0, bin, epilogue, power, out0, n_out_bins, stream);
}

extern "C" int rvb_stft_bin_folded_f16(const void* a_hi, const void* a_lo, const float* row_scale_inv, int n_seg,
                                       int n_frames, int n_fft, const float* wc_row, const float* ws_row,
                                       const float* p0, float w0, int bin, int epilogue, float power, float* out0,
                                       int n_out_bins, rvb_stream_t stream) {
  RVB_REQUIRE(row_scale_inv, "rvb_stft_bin_folded_f16: null row_scale_inv");
  return launch_bin_folded<__half>("rvb_stft_bin_folded_f16", static_cast<const __half*>(a_hi),
                                   static_cast<const __half*>(a_lo), row_scale_inv, n_seg, n_frames, n_fft, wc_row,
                                   ws_row, p0, w0, bin, epilogue, power, out0, n_out_bins, stream);
}
