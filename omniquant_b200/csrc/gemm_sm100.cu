// K1/K2/K3/K5 contraction: persistent, warp-specialised bf16 GEMM on the 5th-gen
// tensor cores (tcgen05.mma, fp32 accumulators in TMEM, operands staged by TMA
// with 128-byte swizzle), sm_100a only.
//
//   C[b] = alpha * A[b] * B[b]^T (+ bias[n]) (+ residual[b])
//
// Serves QuantLinear's F.linear forward / dgrad / wgrad (reference
// quantize/int_linear.py:62 and its autograd) and QuantMatMul's QK^T / PV and
// their gradients (quantize/int_matmul.py:41-43; models/int_llama_layer.py:143,163)
// on the already fake-quantised bf16 operands.
//
// Operand layouts (element strides; both may be "K-major" or "MN-major" so that
// forward, dgrad and wgrad all read the same buffers without transposes):
//   A: a_mn=0  A[m, k] at A + m*lda + k      a_mn=1  A[m, k] at A + k*lda + m
//   B: b_mn=0  B[n, k] at B + n*ldb + k      b_mn=1  B[n, k] at B + k*ldb + n
// Tile 128 x BN x 64 (BN = 256 or 128), one CTA per SM looping over tiles,
// warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4-7 = epilogue (TMEM -> registers -> global), two TMEM accumulator
// stages so the epilogue of tile i overlaps the main loop of tile i+1.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"
#include "mbar.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 256;

struct GemmParams {
  int M, N, K, batch;
  int a_mn, b_mn;
  void* C;
  int c_dtype;  // 0 bf16, 1 fp32, 2 fp16
  long ldc, strideC;
  const float* bias;
  const bf16* res;
  long ldr, strideR;
  float alpha;
  int num_m, num_n;
  int a_bcast, b_bcast;  // operand shared by all batch entries (GQA: one K/V head serves several Q heads)
  int accumulate;        // fp32 output only: C += result
  int causal;  // 0 none; 1 skip tiles with n0 > m0+BM-1; 2 k < m0+BM; 3 k >= m0 (rounded to BK)
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// same instruction with the descriptors given as (low word, shared high word)
__device__ __forceinline__ void tc_mma_bf16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int BN>
struct Cfg {
  static constexpr int STAGES = BN == 256 ? 4 : 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void tile_coords(const GemmParams& p, int tile, int& m0, int& n0, int& b) {
  const int mb = tile % p.num_m;
  const int rest = tile / p.num_m;
  m0 = mb * BM;
  n0 = (rest % p.num_n);
  b = rest / p.num_n;
}
__device__ __forceinline__ bool tile_skipped(const GemmParams& p, int m0, int n0) { return p.causal == 1 && n0 > m0 + BM - 1; }
__device__ __forceinline__ void k_range(const GemmParams& p, int m0, int& kb0, int& kb1) {
  kb0 = 0;
  kb1 = (p.K + BK - 1) / BK;
  if (p.causal == 2) kb1 = min(kb1, (min(m0 + BM, p.K) + BK - 1) / BK);
  if (p.causal == 3) kb0 = min(m0 / BK, kb1);
}

// one 32-column chunk of an accumulator row: alpha, bias, residual, optional fp32 accumulate, store as bf16 / fp32 / fp16
__device__ __forceinline__ void store_chunk(const GemmParams& p, const uint32_t (&r)[32], int row, int b, int nc, const bf16* resr,
                                            bool vec_ok) {
  {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
    const int valid = min(32, p.N - nc);
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < valid) v[j] += p.bias[nc + j];
  }
  if (resr) {
    if (valid == 32 && p.ldr % 8 == 0 && p.strideR % 8 == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float t[8];
        load_bf8(resr + nc + j, t);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[j + e] += t[e];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < valid) v[j] += bf2f(resr[nc + j]);
    }
  }
  if (p.c_dtype == 1) {
    float* cr = reinterpret_cast<float*>(p.C) + (long)b * p.strideC + (long)row * p.ldc + nc;
    if (p.accumulate) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < valid) v[j] += cr[j];
    }
    if (valid == 32 && vec_ok) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(cr + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < valid) cr[j] = v[j];
    }
  } else if (p.c_dtype == 2) {
    __half* cr = reinterpret_cast<__half*>(p.C) + (long)b * p.strideC + (long)row * p.ldc + nc;
    if (valid == 32 && vec_ok) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 u;
        __half2 h0 = __floats2half2_rn(v[j], v[j + 1]), h1 = __floats2half2_rn(v[j + 2], v[j + 3]);
        __half2 h2 = __floats2half2_rn(v[j + 4], v[j + 5]), h3 = __floats2half2_rn(v[j + 6], v[j + 7]);
        u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
        u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(cr + j) = u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < valid) cr[j] = __float2half_rn(v[j]);
    }
  } else {
    bf16* cr = reinterpret_cast<bf16*>(p.C) + (long)b * p.strideC + (long)row * p.ldc + nc;
    if (valid == 32 && vec_ok) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float t[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = v[j + e];
        store_bf8(cr + j, t);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < valid) cr[j] = f2bf(v[j]);
    }
  }
}
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;                    // [STAGES]
  uint64_t* empty = bars + C::STAGES;       // [STAGES]
  uint64_t* tfull = bars + 2 * C::STAGES;   // [2]
  uint64_t* tempty = tfull + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.num_m * p.num_n * p.batch;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m0, nb, b;
        tile_coords(p, tile, m0, nb, b);
        const int n0 = nb * BN;
        if (tile_skipped(p, m0, n0)) continue;
        int kb0, kb1;
        k_range(p, m0, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          mbar_expect_tx(&full[stage], C::STAGE_BYTES);
          const int k0 = kb * BK;
          if (!p.a_mn) {
            tma_load_3d(sa, &tmA, &full[stage], k0, m0, p.a_bcast ? 0 : b);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_3d(sa + j * (64 * BK * 2), &tmA, &full[stage], m0 + 64 * j, k0, p.a_bcast ? 0 : b);
          }
          if (!p.b_mn) {
            tma_load_3d(sb, &tmB, &full[stage], k0, n0, p.b_bcast ? 0 : b);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_3d(sb + j * (64 * BK * 2), &tmB, &full[stage], n0 + 64 * j, k0, p.b_bcast ? 0 : b);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, majors, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.a_mn ? 1 : 0) << 15) |
                             ((uint32_t)(p.b_mn ? 1 : 0) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      // K-major: rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused (=16 B).
      // MN-major: 64-element (128 B) MN blocks 64*BK*2 B apart (LBO), 8-k-row groups 1024 B apart (SBO).
      const uint32_t a_lbo = p.a_mn ? 64 * BK * 2 : 16, b_lbo = p.b_mn ? 64 * BK * 2 : 16;
      // The descriptor's high word is a constant and a k-step / ring stage only adds a multiple of 16 bytes to the start-address
      // field of the low word: everything else is hoisted out of the issue loop (the issuing thread is a serial resource:
      // one MMA must be issued every 128 cycles to keep the tensor pipe busy).
      const uint32_t a_k16 = (p.a_mn ? UMMA_K * 128 : UMMA_K * 2) >> 4, b_k16 = (p.b_mn ? UMMA_K * 128 : UMMA_K * 2) >> 4;
      const uint32_t sa0 = __shfl_sync(0x1u, smem_u32(smem), 0);
      const uint32_t alo0 = ((sa0 & 0x3FFFF) >> 4) | (((a_lbo >> 4) & 0x3FFF) << 16);
      const uint32_t blo0 = (((sa0 + C::A_BYTES) & 0x3FFFF) >> 4) | (((b_lbo >> 4) & 0x3FFF) << 16);
      constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m0, nb, b;
        tile_coords(p, tile, m0, nb, b);
        if (tile_skipped(p, m0, nb * BN)) continue;
        int kb0, kb1;
        k_range(p, m0, kb0, kb1);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t alo = alo0 + (uint32_t)stage * (C::STAGE_BYTES >> 4), blo = blo0 + (uint32_t)stage * (C::STAGE_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            tc_mma_bf16_lo(d_tmem, alo + k * a_k16, blo + k * b_k16, DESC_HI, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          tc_commit(&empty[stage]);  // frees the smem stage when these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull[acc]);  // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = p.c_dtype == 1 ? (p.ldc % 4 == 0 && p.strideC % 4 == 0) : (p.ldc % 8 == 0 && p.strideC % 8 == 0);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int m0, nb, b;
      tile_coords(p, tile, m0, nb, b);
      const int n0 = nb * BN;
      if (tile_skipped(p, m0, n0)) continue;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      const bf16* resr = p.res ? p.res + (long)b * p.strideR + (long)row * p.ldr : nullptr;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tc_ld32(taddr + c0, r);
        tc_wait_ld();
        if (row < p.M) store_chunk(p, r, row, b, n0 + c0, resr, vec_ok);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

// ================================================================== CTA-pair variant (cta_group::2)
// Two CTAs on the two SMs of a TPC (a 2-CTA cluster) compute one 256 x 256 output tile: CTA r of the pair owns rows
// [128 r, 128 r + 128) of the tile (its A operand and its accumulator: 128 TMEM lanes x 256 columns) and stages HALF of the B
// tile (N rows [128 r, 128 r + 128)); tcgen05.mma.cta_group::2 issued by the leader (rank 0) reads A from each CTA's own shared
// memory and the two B halves from both.  Per CTA and k-block that is 16 KB of A + 16 KB of B for 2 x 128 x 256 x 64 flops --
// two thirds of the L2 -> SM operand traffic of the single-CTA 128 x 256 tile (48 KB), which is what bounds that kernel
// (its 96 B/clk/SM at the nominal MMA rate is more than the L2 delivers), and six stages fit instead of four.
//   barriers: full[s] lives in the LEADER (count 1: the leader's arrive.expect_tx for both CTAs' bytes; the peer's TMA
//   completes its bytes on the leader's barrier -- address through mapa); empty[s], tfull[a] in both CTAs, signalled
//   by the leader's tcgen05.commit ... multicast::cluster; tempty[a] in the leader, count 8 (the four epilogue warps of each CTA).
// Not batched, no causal modes (the linear layers of the block; attention has its own kernels).
// shared::cluster address of the same object in the pair's leader CTA (rank 0)
__device__ __forceinline__ uint32_t leader_addr(const void* ptr) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(ptr)), "r"(0));
  return r;
}
constexpr int STAGES2 = 6;
constexpr int A2_BYTES = BM * BK * 2, B2_BYTES = 128 * BK * 2, STAGE2_BYTES = A2_BYTES + B2_BYTES;
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 1024 + 256;

__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(leader_addr(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {     // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2sm(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {   // release at cluster scope on the leader's copy
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(leader_addr(bar)) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Tile order of the CTA-pair kernel: m fastest inside groups of PAIR_GM m-tiles, n sweeping the group, group after group.
// A wave of 74 pair tiles then touches 8 A tiles and ~9 B tiles instead of 74 A tiles and one B tile: with plain m-fastest
// order the wgrad of gate|up (86 x 16 tiles, A = dY^T 90 MB) re-read A for every n tile -- 675 MB of DRAM reads for 107 MB
// of operands (ncu, profiles/r2_gemm_ncu_full_summary.csv).  Products with at most PAIR_GM m-tiles are unchanged.
constexpr int PAIR_GM = 8;
__device__ __forceinline__ void pair_tile(const GemmParams& p, int tile, int& mi, int& ni) {
  const int per_group = PAIR_GM * p.num_n;
  const int grp = tile / per_group, local = tile - grp * per_group;
  const int gm = min(PAIR_GM, p.num_m - grp * PAIR_GM);
  ni = local / gm;
  mi = grp * PAIR_GM + (local - ni * gm);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  constexpr int BN = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES2 * STAGE2_BYTES);
  uint64_t* full = bars;                  // [STAGES2]  (used in the leader)
  uint64_t* empty = bars + STAGES2;       // [STAGES2]
  uint64_t* tfull = bars + 2 * STAGES2;   // [2]
  uint64_t* tempty = tfull + 2;           // [2]        (used in the leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int total_tiles = p.num_m * p.num_n;      // num_m counts 256-row tiles here

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);  // one arrive per epilogue warp of either CTA
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();      // barriers initialised and TMEM allocated in both CTAs before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int kb1 = (p.K + BK - 1) / BK;
      for (int tile = pair; tile < total_tiles; tile += num_pairs) {
        int mi, ni;
        pair_tile(p, tile, mi, ni);
        const int m0 = mi * 256 + 128 * (int)rank;      // this CTA's 128 rows
        const int n0 = ni * BN + 128 * (int)rank;       // this CTA's half of the B tile
        for (int kb = 0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE2_BYTES;
          uint8_t* sb = sa + A2_BYTES;
          if (rank == 0) mbar_expect_tx(&full[stage], 2 * STAGE2_BYTES);
          const int k0 = kb * BK;
          if (!p.a_mn) {
            tma_load_3d_2sm(sa, &tmA, &full[stage], k0, m0, 0);
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_3d_2sm(sa + j * (64 * BK * 2), &tmA, &full[stage], m0 + 64 * j, k0, 0);
          }
          if (!p.b_mn) {
            tma_load_3d_2sm(sb, &tmB, &full[stage], k0, n0, 0);
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_3d_2sm(sb + j * (64 * BK * 2), &tmB, &full[stage], n0 + 64 * j, k0, 0);
          }
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader only) =====================
    if (lane == 0 && rank == 0) {
      // M = 256 over the pair, N = 256 (each CTA contributes 128 rows of B)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.a_mn ? 1 : 0) << 15) |
                             ((uint32_t)(p.b_mn ? 1 : 0) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t a_lbo = p.a_mn ? 64 * BK * 2 : 16, b_lbo = p.b_mn ? 64 * BK * 2 : 16;
      const uint32_t a_k16 = (p.a_mn ? UMMA_K * 128 : UMMA_K * 2) >> 4, b_k16 = (p.b_mn ? UMMA_K * 128 : UMMA_K * 2) >> 4;
      const uint32_t sa0 = __shfl_sync(0x1u, smem_u32(smem), 0);
      const uint32_t alo0 = ((sa0 & 0x3FFFF) >> 4) | (((a_lbo >> 4) & 0x3FFF) << 16);
      const uint32_t blo0 = (((sa0 + A2_BYTES) & 0x3FFFF) >> 4) | (((b_lbo >> 4) & 0x3FFF) << 16);
      constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const int kb1 = (p.K + BK - 1) / BK;
      for (int tile = pair; tile < total_tiles; tile += num_pairs) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t alo = alo0 + (uint32_t)stage * (STAGE2_BYTES >> 4), blo = blo0 + (uint32_t)stage * (STAGE2_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            tc_mma_bf16_2sm(d_tmem, alo + k * a_k16, blo + k * b_k16, DESC_HI, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit_2sm(&empty[stage]);  // frees the stage in both CTAs when these MMAs retire
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
        tc_commit_2sm(&tfull[acc]);  // accumulator complete: both CTAs' epilogues
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, each its own 128 rows) =====================
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = p.c_dtype == 1 ? (p.ldc % 4 == 0) : (p.ldc % 8 == 0);
    for (int tile = pair; tile < total_tiles; tile += num_pairs) {
      int mi, ni;
      pair_tile(p, tile, mi, ni);
      const int m0 = mi * 256 + 128 * (int)rank;
      const int n0 = ni * BN;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      const bf16* resr = p.res ? p.res + (long)row * p.ldr : nullptr;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tc_ld32(taddr + c0, r);
        tc_wait_ld();
        if (row < p.M) store_chunk(p, r, row, 0, n0 + c0, resr, vec_ok);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();      // nobody frees TMEM or exits while the peer may still signal or read across the pair
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 3-D bf16 tensor map: dims (inner, rows, batch) with element strides (1, ld, bstride); box (64, box_rows, 1)
int make_map(CUtensorMap* map, const void* base, long inner, long rows, long batch, long ld, long bstride, int box_rows) {
  EncodeTiledFn enc = get_encode();
  OQ_REQUIRE(enc, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batch > 1 ? bstride : ld * rows) * 2};
  if (strides[1] == 0) strides[1] = strides[0];
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  OQ_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): base=%p inner=%ld rows=%ld batch=%ld ld=%ld bstride=%ld",
             (int)r, base, inner, rows, batch, ld, bstride);
  return 0;
}

template <int BN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, GemmParams p, int max_ctas, cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) {
      oq_set_error("oq_gemm_bf16: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  p.num_m = (p.M + BM - 1) / BM;
  p.num_n = (p.N + BN - 1) / BN;
  const long tiles = (long)p.num_m * p.num_n * p.batch;
  int grid = (int)std::min<long>(tiles, oq_num_sms());
  if (max_ctas > 0) grid = std::min(grid, max_ctas);
  gemm_kernel<BN><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, p);
  return oq_check_launch("oq_gemm_bf16");
}

int launch2(const CUtensorMap& ta, const CUtensorMap& tb, GemmParams p, int max_ctas, cudaStream_t stream) {
  static int max_pairs = 0;
  if (!max_pairs) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES);
    if (e != cudaSuccess) {
      oq_set_error("oq_gemm_bf16: cudaFuncSetAttribute (CTA-pair kernel) failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * oq_num_sms());
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM2_BYTES;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, gemm2_kernel, &cfg);     // cluster size comes from __cluster_dims__
    if (e != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = oq_num_sms() / 2;
    }
    max_pairs = std::min(n, oq_num_sms() / 2);
  }
  p.num_m = (p.M + 255) / 256;
  p.num_n = (p.N + 255) / 256;
  const long tiles = (long)p.num_m * p.num_n;
  int pairs = (int)std::min<long>(tiles, max_pairs);
  if (max_ctas > 1) pairs = std::min(pairs, max_ctas / 2);
  gemm2_kernel<<<2 * pairs, NUM_THREADS, SMEM2_BYTES, stream>>>(ta, tb, p);
  return oq_check_launch("oq_gemm_bf16 (CTA pair)");
}

// the CTA-pair kernel serves plain (unbatched, non-causal) products with at least one full pair tile per SM pair
bool use_pair_kernel(int M, int N, int batch, int causal, int tile_n) {
  static const int mode = [] { const char* e = getenv("OQ_GEMM_2CTA"); return e ? atoi(e) : 1; }();
  if (!mode || batch != 1 || causal || tile_n == 128) return false;
  const long tiles = (long)((M + 255) / 256) * ((N + 255) / 256);
  return M >= 256 && N >= 256 && tiles >= oq_num_sms() / 2;
}

}  // namespace

extern "C" int oq_gemm_bf16(const void* A, int a_mn, long lda, long strideA, const void* B, int b_mn, long ldb,
                            long strideB, void* C, int c_dtype, long ldc, long strideC, int M, int N, int K, int batch,
                            const float* bias, const void* residual, long ldr, long strideR, float alpha, int causal,
                            int accumulate, int tile_n, int max_ctas, cudaStream_t stream) {
  OQ_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && batch > 0, "oq_gemm_bf16: null pointer or bad shape M=%d N=%d K=%d batch=%d",
             M, N, K, batch);
  OQ_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && strideA % 8 == 0 && strideB % 8 == 0,
             "oq_gemm_bf16: operand strides must be multiples of 8 elements (TMA 16-byte rule): lda=%ld ldb=%ld sA=%ld sB=%ld",
             lda, ldb, strideA, strideB);
  OQ_REQUIRE((uintptr_t)A % 16 == 0 && (uintptr_t)B % 16 == 0, "oq_gemm_bf16: operand base pointers must be 16-byte aligned");
  OQ_REQUIRE(causal >= 0 && causal <= 3, "oq_gemm_bf16: bad causal mode %d", causal);
  OQ_REQUIRE(c_dtype >= 0 && c_dtype <= 2, "oq_gemm_bf16: c_dtype must be 0 (bf16), 1 (fp32) or 2 (fp16)");
  OQ_REQUIRE(tile_n == 0 || tile_n == 128 || tile_n == 256, "oq_gemm_bf16: tile_n must be 0, 128 or 256");
  const bool pair = use_pair_kernel(M, N, batch, causal, tile_n);
  if (pair) tile_n = 256;
  if (tile_n == 0) {
    // BN=256 halves the A re-reads; use BN=128 when N is small or the wider tile leaves SMs idle
    const long t256 = (long)((M + BM - 1) / BM) * ((N + 255) / 256) * batch;
    tile_n = (N <= 128 || t256 < oq_num_sms()) ? 128 : 256;
  }
  GemmParams p = {};
  p.M = M; p.N = N; p.K = K; p.batch = batch; p.a_mn = a_mn ? 1 : 0; p.b_mn = b_mn ? 1 : 0;
  p.C = C; p.c_dtype = c_dtype; p.ldc = ldc; p.strideC = strideC; p.bias = bias; p.res = (const bf16*)residual;
  p.ldr = ldr; p.strideR = strideR; p.alpha = alpha; p.causal = causal;
  OQ_REQUIRE(!accumulate || c_dtype == 1, "oq_gemm_bf16: accumulate needs an fp32 output");
  p.accumulate = accumulate ? 1 : 0;
  p.a_bcast = (batch > 1 && strideA == 0) ? 1 : 0;   // stride 0 = the operand is shared by every batch entry
  p.b_bcast = (batch > 1 && strideB == 0) ? 1 : 0;
  const int batchA = p.a_bcast ? 1 : batch, batchB = p.b_bcast ? 1 : batch;
  CUtensorMap ta, tb;
  // K-major: inner = K, rows = M|N, box rows = BM|BN.   MN-major: inner = M|N, rows = K, box = 64 x 64.
  if (!p.a_mn) { if (make_map(&ta, A, K, M, batchA, lda, strideA, BM)) return -1; }
  else         { if (make_map(&ta, A, M, K, batchA, lda, strideA, BK)) return -1; }
  if (!p.b_mn) { if (make_map(&tb, B, K, N, batchB, ldb, strideB, pair ? 128 : tile_n)) return -1; }
  else         { if (make_map(&tb, B, N, K, batchB, ldb, strideB, BK)) return -1; }
  if (pair) return launch2(ta, tb, p, max_ctas, stream);
  return tile_n == 256 ? launch<256>(ta, tb, p, max_ctas, stream) : launch<128>(ta, tb, p, max_ctas, stream);
}
