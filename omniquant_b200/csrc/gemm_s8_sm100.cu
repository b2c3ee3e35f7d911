// Exact-integer forward contraction of QuantLinear for configurations that quantise BOTH operands
// (W4A4 / W6A6): the operands are the signed integer codes (code - zero point) of the fake-quantised activations and
// weights, the sum over K is exact (tcgen05.mma kind::i8 = UTCIMMA, s32 accumulators in TMEM), the per-token and
// per-channel scales are applied in the epilogue:
//
//   C[m,n] = alpha * ( rs[m]*cs[n] * sum_k A[m,k]*B[n,k]
//                      (+ rt[m]*cbeta[n] + rbeta[m]*ct[n] + K*rbeta[m]*cbeta[n]) )  (+ bias[n]) (+ residual[m,n])
//
// Reference: QuantLinear.forward, quantize/int_linear.py:48-65 (F.linear on the two fake-quantised tensors):
// x_q[m,k] = (c_a - z_a[m])*s_a[m], w_q[n,k] = (c_w - z_w[n])*s_w[n]  (quantize/quantizer.py:93-100), so
// F.linear(x_q, w_q)[m,n] = s_a[m]*s_w[n]*sum_k (c_a - z_a)(c_w - z_w).  With |c - z| <= 127 and K*127^2 < 2^31 the integer
// sum has no rounding at all, unlike the bf16 path which rounds every dequantised operand to 8 mantissa bits.  A zero point
// outside the code range (a row that does not straddle zero) is split as z = z0 + (z - z0) with z0 clamped into range: the
// caller passes beta = (z0 - z)*s and t = s*sum_k(c - z0) per row / column and the epilogue adds the cross terms.
//
// Used by the block engine for the four forward linears of W4A4 / W6A6 runs (engine.py: _lin_fwd); the operands are the s8
// side outputs of the activation / weight quantiser kernels (common.cuh S8Side).  Validated bit for bit against the same
// expression evaluated by torch (tests/test_kernels_gpu.py::test_gemm_s8_exact).  Same pipeline as gemm_sm100.cu:
// warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue, two TMEM accumulator stages.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"
#include "mbar.cuh"

namespace {

constexpr int BM = 128;
constexpr int BKB = 128;   // bytes (= s8 elements) per smem row: one 128-byte swizzle row
constexpr int UMMA_KB = 32;  // bytes per MMA k-step (32 s8 elements)
constexpr int NUM_THREADS = 256;

struct S8Params {
  int M, N, K;
  void* C;
  int c_dtype;  // 0 bf16, 1 fp32, 2 fp16
  long ldc;
  const float* bias;
  const bf16* res;
  long ldr;
  const float* rs; const float* cs;
  const float* rt; const float* rbeta; const float* ct; const float* cbeta;
  float alpha;
  int num_m, num_n;
};

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_s8_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- CTA-pair (cta_group::2) variants: see gemm_sm100.cu (gemm2_kernel) for the protocol
__device__ __forceinline__ uint32_t leader_addr(const void* ptr) {      // same object in the pair's rank-0 CTA
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(ptr)), "r"(0));
  return r;
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(leader_addr(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc_mma_s8_2sm(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(leader_addr(bar)) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
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
__device__ __forceinline__ float f4c(const float4& v, int j) { return reinterpret_cast<const float*>(&v)[j]; }

// PAIR: two CTAs (a 2-CTA cluster) per 256 x 256 tile; each stages its 128 rows of A and 128 of the 256 rows of B
template <int BN, bool PAIR>
struct Cfg {
  static constexpr int STAGES = (PAIR || BN != 256) ? 6 : 4;
  static constexpr int A_BYTES = BM * BKB;
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BKB;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int COLV_BYTES = 2 * 4 * BN * 4;   // per accumulator stage: column scale | beta | t | bias of the tile
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + COLV_BYTES;
};

template <int BN, bool PAIR>
__device__ __forceinline__ void gemm_s8_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const S8Params& p) {
  using C = Cfg<BN, PAIR>;
  constexpr int TM = PAIR ? 2 * BM : BM;        // rows of one scheduled tile
  uint32_t rank = 0;
  if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int first_tile = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int row_off = PAIR ? BM * (int)rank : 0, bcol_off = PAIR ? (BN / 2) * (int)rank : 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* tfull = bars + 2 * C::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* colv = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.num_m * p.num_n;
  const int num_kb = (p.K + BKB - 1) / BKB;

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
      mbar_init(&tempty[i], PAIR ? 8 : 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int m0 = (tile % p.num_m) * TM + row_off, n0 = (tile / p.num_m) * BN + bcol_off;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          if (PAIR) {      // the leader's barrier collects the bytes of both CTAs
            if (rank == 0) mbar_expect_tx(&full[stage], 2 * C::STAGE_BYTES);
            tma_load_2d_2sm(sa, &tmA, &full[stage], kb * BKB, m0);
            tma_load_2d_2sm(sa + C::A_BYTES, &tmB, &full[stage], kb * BKB, n0);
          } else {
            mbar_expect_tx(&full[stage], C::STAGE_BYTES);
            tma_load_2d(sa, &tmA, &full[stage], kb * BKB, m0);           // out-of-range rows / K tail are zero-filled
            tma_load_2d(sa + C::A_BYTES, &tmB, &full[stage], kb * BKB, n0);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // instruction descriptor: D = s32 (2), A = B = int8 (1), both K-major, N>>3, M>>4 (M = 256 over a CTA pair)
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      // descriptor words hoisted out of the issue loop: the high word is a constant, a stage / k-step only adds to the
      // start-address field of the low word.  An int8 MMA of this shape occupies the tensor pipe for 64 cycles only, so
      // the issuing thread has to stay well below that per instruction.
      const uint32_t sa0 = __shfl_sync(0x1u, smem_u32(smem), 0);
      const uint32_t alo0 = ((sa0 & 0x3FFFF) >> 4) | (1u << 16), blo0 = (((sa0 + C::A_BYTES) & 0x3FFFF) >> 4) | (1u << 16);
      constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t alo = alo0 + (uint32_t)stage * (C::STAGE_BYTES >> 4), blo = blo0 + (uint32_t)stage * (C::STAGE_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BKB / UMMA_KB; ++k) {
            if (PAIR) tc_mma_s8_2sm(d_tmem, alo + k * (UMMA_KB >> 4), blo + k * (UMMA_KB >> 4), DESC_HI, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            else tc_mma_s8_lo(d_tmem, alo + k * (UMMA_KB >> 4), blo + k * (UMMA_KB >> 4), DESC_HI, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          if (PAIR) tc_commit_2sm(&empty[stage]);
          else tc_commit(&empty[stage]);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if (PAIR) tc_commit_2sm(&tfull[acc]);
        else tc_commit(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    const float fK = (float)p.K;
    // The per-column epilogue vectors (scale, beta, t, bias) are the same for every row: staged once per tile in shared
    // memory by the 128 epilogue threads (before the accumulator is awaited, so it overlaps the main loop) and read back as
    // broadcast 16-byte loads.  Reading them from global memory per element made the epilogue LSU-bound (4 loads per
    // output: ~11k cycles per tile that the K = 4096 main loop of 16k cycles could not hide).
    const int et = (warp - 4) * 32 + lane;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
      const int m0 = (tile % p.num_m) * TM + row_off, n0 = (tile / p.num_m) * BN;
      float* cv = colv + acc * 4 * BN;
      for (int x = et; x < BN; x += 128) {
        const int c = min(n0 + x, p.N - 1);      // clamped column: loads stay in bounds, the value is not stored
        cv[x] = p.cs[c];
        cv[BN + x] = p.rt ? p.cbeta[c] : 0.f;
        cv[2 * BN + x] = p.rt ? p.ct[c] : 0.f;
        cv[3 * BN + x] = p.bias ? p.bias[c] : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      const bool live = row < p.M;
      const float rs = live ? p.rs[row] : 0.f;
      const float rt = (live && p.rt) ? p.rt[row] : 0.f, rb = (live && p.rt) ? p.rbeta[row] : 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tc_ld32(taddr + c0, r);
        tc_wait_ld();
        if (live) {
          const int nc = n0 + c0;
          const int valid = min(32, p.N - nc);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 cs4 = *reinterpret_cast<const float4*>(cv + c0 + j);
            float y[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) y[e] = __fmul_rn(__fmul_rn(__int2float_rn((int)r[j + e]), rs), f4c(cs4, e));
            if (p.rt) {
              const float4 cb4 = *reinterpret_cast<const float4*>(cv + BN + c0 + j), ct4 = *reinterpret_cast<const float4*>(cv + 2 * BN + c0 + j);
#pragma unroll
              for (int e = 0; e < 4; ++e) y[e] += rt * f4c(cb4, e) + rb * f4c(ct4, e) + fK * rb * f4c(cb4, e);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) y[e] *= p.alpha;
            if (p.bias) {
              const float4 b4 = *reinterpret_cast<const float4*>(cv + 3 * BN + c0 + j);
#pragma unroll
              for (int e = 0; e < 4; ++e) y[e] += f4c(b4, e);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) v[j + e] = y[e];
          }
          if (p.res) {
            const bf16* resr = p.res + (long)row * p.ldr + nc;
            if (valid == 32 && p.ldr % 8 == 0 && ((uintptr_t)p.res % 16) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                float t[8];
                load_bf8(resr + j, t);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[j + e] += t[e];
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < valid) v[j] += bf2f(resr[j]);
            }
          }
          if (p.c_dtype == 1) {
            float* cr = reinterpret_cast<float*>(p.C) + (long)row * p.ldc + nc;
            if (valid == 32 && p.ldc % 4 == 0 && ((uintptr_t)p.C % 16) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(cr + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < valid) cr[j] = v[j];
            }
          } else if (p.c_dtype == 2) {
            __half* cr = reinterpret_cast<__half*>(p.C) + (long)row * p.ldc + nc;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < valid) cr[j] = __float2half_rn(v[j]);
          } else {
            bf16* cr = reinterpret_cast<bf16*>(p.C) + (long)row * p.ldc + nc;
            if (valid == 32 && p.ldc % 8 == 0 && ((uintptr_t)p.C % 16) == 0) {
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(&tempty[acc]);
        else mbar_arrive(&tempty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_s8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const S8Params p) {
  gemm_s8_body<BN, false>(tmA, tmB, p);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_s8_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const S8Params p) {
  gemm_s8_body<256, true>(tmA, tmB, p);
}

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

// 2-D s8 tensor map: dims (K, rows), row stride ld bytes; box (128, box_rows)
int make_map_s8(CUtensorMap* map, const void* base, long K, long rows, long ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  OQ_REQUIRE(enc, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld};
  cuuint32_t box[2] = {(cuuint32_t)BKB, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  OQ_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (s8) failed (%d): base=%p K=%ld rows=%ld ld=%ld", (int)r, base, K, rows, ld);
  return 0;
}

int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, S8Params p, cudaStream_t stream) {
  using C = Cfg<256, true>;
  static int max_pairs = 0;
  if (!max_pairs) {
    cudaError_t e = cudaFuncSetAttribute(gemm_s8_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) {
      oq_set_error("oq_gemm_s8: cudaFuncSetAttribute (CTA-pair kernel) failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * oq_num_sms());
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, gemm_s8_pair_kernel, &cfg);
    if (e != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = oq_num_sms() / 2;
    }
    max_pairs = std::min(n, oq_num_sms() / 2);
  }
  p.num_m = (p.M + 2 * BM - 1) / (2 * BM);
  p.num_n = (p.N + 255) / 256;
  const long tiles = (long)p.num_m * p.num_n;
  const int pairs = (int)std::min<long>(tiles, max_pairs);
  gemm_s8_pair_kernel<<<2 * pairs, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, p);
  return oq_check_launch("oq_gemm_s8 (CTA pair)");
}

template <int BN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, S8Params p, cudaStream_t stream) {
  using C = Cfg<BN, false>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_s8_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) {
      oq_set_error("oq_gemm_s8: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  p.num_m = (p.M + BM - 1) / BM;
  p.num_n = (p.N + BN - 1) / BN;
  const long tiles = (long)p.num_m * p.num_n;
  const int grid = (int)std::min<long>(tiles, oq_num_sms());
  gemm_s8_kernel<BN><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, p);
  return oq_check_launch("oq_gemm_s8");
}

}  // namespace

extern "C" int oq_gemm_s8(const void* A, long lda, const void* B, long ldb, void* C, int c_dtype, long ldc, int M, int N, int K,
                          const float* row_scale, const float* col_scale, const float* row_t, const float* row_beta,
                          const float* col_t, const float* col_beta, float alpha, const float* bias, const void* residual,
                          long ldr, cudaStream_t stream) {
  OQ_REQUIRE(A && B && C && row_scale && col_scale && M > 0 && N > 0 && K > 0, "oq_gemm_s8: null pointer or bad shape");
  OQ_REQUIRE(lda % 16 == 0 && ldb % 16 == 0 && (uintptr_t)A % 16 == 0 && (uintptr_t)B % 16 == 0,
             "oq_gemm_s8: operand strides and base pointers must be multiples of 16 bytes (TMA)");
  OQ_REQUIRE(c_dtype >= 0 && c_dtype <= 2, "oq_gemm_s8: c_dtype must be 0 (bf16), 1 (fp32) or 2 (fp16)");
  const bool corr = row_t || row_beta || col_t || col_beta;
  OQ_REQUIRE(!corr || (row_t && row_beta && col_t && col_beta), "oq_gemm_s8: the zero-point correction needs all four vectors");
  OQ_REQUIRE((long)K * 127 * 127 < (1L << 31), "oq_gemm_s8: K too large for an exact s32 sum");
  S8Params p = {};
  p.M = M; p.N = N; p.K = K; p.C = C; p.c_dtype = c_dtype; p.ldc = ldc; p.bias = bias;
  p.res = (const bf16*)residual; p.ldr = ldr;
  p.alpha = alpha;
  p.rs = row_scale; p.cs = col_scale; p.rt = row_t; p.rbeta = row_beta; p.ct = col_t; p.cbeta = col_beta;
  const long t256 = (long)((M + BM - 1) / BM) * ((N + 255) / 256);
  const int tile_n = (N <= 128 || t256 < oq_num_sms()) ? 128 : 256;
  static const int pair_mode = [] { const char* e = getenv("OQ_GEMM_2CTA"); return e ? atoi(e) : 1; }();
  const long tpair = (long)((M + 2 * BM - 1) / (2 * BM)) * ((N + 255) / 256);
  const bool pair = pair_mode && M >= 2 * BM && N >= 256 && tpair >= oq_num_sms() / 2;
  CUtensorMap ta, tb;
  if (make_map_s8(&ta, A, K, M, lda, BM)) return -1;
  if (make_map_s8(&tb, B, K, N, ldb, pair ? 128 : tile_n)) return -1;
  if (pair) return launch_pair(ta, tb, p, stream);
  return tile_n == 256 ? launch<256>(ta, tb, p, stream) : launch<128>(ta, tb, p, stream);
}
