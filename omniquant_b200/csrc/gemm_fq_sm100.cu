// K1 (BASELINE.json north_star, first subsystem): QuantLinear forward with the LWC weight fake-quantiser fused INTO the
// GEMM -- the fp32 master weight tile comes in by TMA, is clamped / quantised / rounded / dequantised by transform warps
// in shared memory, and the 5th-gen tensor cores consume the bf16 result straight from there; the fake-quantised weight
// never exists in HBM.
//
//   C[m, n] = sum_k A[m, k] * fq(W)[n, k] (+ bias[n]) (+ residual[m, n])
//   fq(W)[n, k] = bf16( (clamp(rint(W[n, k] / s[n, g]) + z[n, g], 0, 2^bits - 1) - z[n, g]) * s[n, g] ),  g = k / group
//
// Reference: QuantLinear.forward with use_weight_quant (quantize/int_linear.py:48-65) on the weight produced by
// UniformAffineQuantizer.fake_quant (quantize/quantizer.py:84-105); the per-row / per-group scale and zero point are the
// LWC-clipped ones of :122-147 (oq_lwc_qparams below evaluates them from the group extrema, which are constant per block
// for LWC-only configurations).  The operand the MMA sees is bit-identical to the bf16 output of oq_weight_quant_fwd
// (tests/test_kernels_gpu.py::test_gemm_fq_operand_is_bit_exact).
//
// This kernel exists to MEASURE the fused formulation against the pre-pass + GEMM the engine uses (DESIGN.md section 6;
// profiles/r2_k1_ab.txt): quantising a 128 x 64 fp32 tile costs ~26 issue slots per row per warp (exact-rounding guard
// included), i.e. ~830 cycles per k-block with 8 transform warps, against 512 cycles of MMA for the two 128-row halves
// of a 256 x 128 output tile that share the transformed operand -- and the same weight tile is transformed again by every
// CTA of its column (M / 256 times), while the pre-pass quantises each weight once per step at HBM speed.
//
// Roles (512 threads): warp 0 TMA producer (A ring 3 x 32 KB, fp32 W ring 2 x 32 KB), warp 1 MMA issuer, warp 2 TMEM
// allocator, warps 4-7 epilogue, warps 8-15 transform (fp32 stage -> swizzled bf16 operand ring 3 x 16 KB).
#include <cuda.h>

#include <algorithm>

#include "common.cuh"
#include "mbar.cuh"

namespace {

constexpr int BM = 256, BN = 128, BK = 64, UMMA_K = 16;
constexpr int NUM_THREADS = 512;
constexpr int SA = 3, SW = 2, SB = 3;
constexpr int A_HALF = 128 * BK * 2;        // one 128-row K-major bf16 block: 16 KB
constexpr int A_BYTES = 2 * A_HALF;         // 32 KB
constexpr int W_BYTES = BN * BK * 4;        // fp32 stage: 32 KB
constexpr int B_BYTES = BN * BK * 2;        // bf16 operand: 16 KB
constexpr int SMEM_BYTES = SA * A_BYTES + SW * W_BYTES + SB * B_BYTES + BN * 16 /*row parameters*/ + 1024 + 256;

struct FqParams {
  int M, N, K;
  const float* scale; const float* zp;   // [N, G]
  int G, group, n_bits;
  bf16* C; long ldc;
  const float* bias;
  const bf16* res; long ldr;
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
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
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
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {     // K-major, SWIZZLE_128B
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct Ring {
  int idx = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void next(int n) {
    if (++idx == n) { idx = 0; phase ^= 1; }
  }
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_fq_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const FqParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                        // [SA][2 halves][128 rows x 128 B]
  uint8_t* sW = sA + SA * A_BYTES;           // [SW][128 rows x 256 B] fp32, not swizzled
  uint8_t* sB = sW + SW * W_BYTES;           // [SB][128 rows x 128 B] bf16, SWIZZLE_128B (written by the transform warps)
  float4* sPar = reinterpret_cast<float4*>(sB + SB * B_BYTES);     // [BN] (scale, zero point, 1/scale, flag) of the tile's rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sPar) + BN * 16);
  uint64_t* a_full = bars;             // [SA]
  uint64_t* a_empty = a_full + SA;     // [SA]
  uint64_t* w_full = a_empty + SA;     // [SW]
  uint64_t* w_empty = w_full + SW;     // [SW]  (8 transform warps arrive)
  uint64_t* b_full = w_empty + SW;     // [SB]  (8 transform warps arrive)
  uint64_t* b_empty = b_full + SB;     // [SB]
  uint64_t* tfull = b_empty + SB;      // [2]
  uint64_t* tempty = tfull + 2;        // [2]   (4 epilogue warps arrive)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.num_m * p.num_n;
  const int num_kb = p.K / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < SW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 8); }
    for (int i = 0; i < SB; ++i) { mbar_init(&b_full[i], 8); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;      // accumulator stage a, half h: columns a * 256 + h * 128

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      Ring ra, rw;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile % p.num_m) * BM, n0 = (tile / p.num_m) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&a_empty[ra.idx], ra.phase ^ 1);
          uint8_t* a = sA + ra.idx * A_BYTES;
          mbar_expect_tx(&a_full[ra.idx], A_BYTES);
          tma_load_2d(a, &tmA, &a_full[ra.idx], kb * BK, m0);
          tma_load_2d(a + A_HALF, &tmA, &a_full[ra.idx], kb * BK, m0 + 128);
          ra.next(SA);
          mbar_wait(&w_empty[rw.idx], rw.phase ^ 1);
          mbar_expect_tx(&w_full[rw.idx], W_BYTES);
          tma_load_2d(sW + rw.idx * W_BYTES, &tmW, &w_full[rw.idx], kb * BK, n0);
          rw.next(SW);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      Ring ra, rb;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&a_full[ra.idx], ra.phase);
          mbar_wait(&b_full[rb.idx], rb.phase);
          tc_fence_after();
          const uint32_t aA = smem_u32(sA + ra.idx * A_BYTES), aB = smem_u32(sB + rb.idx * B_BYTES);
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              tc_mma(tmem + acc * 256 + h * 128, make_desc(aA + h * A_HALF + k * 32), make_desc(aB + k * 32), idesc,
                     (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit(&a_empty[ra.idx]);
          tc_commit(&b_empty[rb.idx]);
          ra.next(SA);
          rb.next(SB);
        }
        tc_commit(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 8) {
    // ===================== transform warps: fp32 stage -> fake-quantised bf16 operand =====================
    const int tw = warp - 8;                 // rows [16 tw, 16 tw + 16) of the tile
    const float qmax = (float)((1 << p.n_bits) - 1);
    float4* par = sPar + tw * 16;            // this warp's 16 rows: (scale, zero point, 1/scale, reciprocal usable)
    Ring rw, rb;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n0 = (tile / p.num_m) * BN;
      const int myrow = n0 + tw * 16 + (lane & 15);
      for (int kb = 0; kb < num_kb; ++kb) {
        // scale / zero point of the warp's rows for the group this k-block lies in: refreshed when the group changes
        // (per-channel: once per tile), kept in shared memory so that a row's parameters are ONE broadcast load
        if (kb == 0 || (p.group && (kb * BK) % p.group == 0)) {
          __syncwarp();
          if (lane < 16) {
            const int gi = p.group ? (kb * BK) / p.group : 0;
            float sc = 1.0f, zp = 0.0f;
            if (myrow < p.N) {
              sc = __ldg(p.scale + (long)myrow * p.G + gi);
              zp = __ldg(p.zp + (long)myrow * p.G + gi);
            }
            const float inv = OQ_DIV(1.0f, sc);
            par[lane] = make_float4(sc, zp, inv, oq_inv_ok(inv) ? 1.0f : 0.0f);
          }
          __syncwarp();
        }
        mbar_wait(&w_full[rw.idx], rw.phase);
        mbar_wait(&b_empty[rb.idx], rb.phase ^ 1);
        const uint8_t* src = sW + rw.idx * W_BYTES;
        uint8_t* dst = sB + rb.idx * B_BYTES;
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
          const int r = tw * 16 + i;
          const float4 pr = par[i];
          const float2 w = *reinterpret_cast<const float2*>(src + r * (BK * 4) + lane * 8);
          OqQParams q;
          q.scale = pr.x; q.zp = pr.y; q.qmax = qmax;
          // branch-free fast path of oq_rint_div (same guard: the reciprocal product decides unless it sits within 2.4e-7
          // relative of a half-integer); ONE rarely taken fallback for the pair -- a branch per element costs a
          // reconvergence barrier each and made the first version of this kernel 6x slower than the pre-pass + GEMM
          const float qa0 = OQ_MUL(w.x, pr.z), qa1 = OQ_MUL(w.y, pr.z);
          const float r0 = rintf(qa0), r1 = rintf(qa1);
          bool redo = pr.w == 0.0f;
          redo |= !(fabsf(fabsf(qa0 - r0) - 0.5f) > fabsf(qa0) * 2.4e-7f + 1e-30f);
          redo |= !(fabsf(fabsf(qa1 - r1) - 0.5f) > fabsf(qa1) * 2.4e-7f + 1e-30f);
          float c0 = fminf(fmaxf(OQ_ADD(r0, pr.y), 0.0f), qmax), c1 = fminf(fmaxf(OQ_ADD(r1, pr.y), 0.0f), qmax);
          if (redo) {      // exact IEEE division (and the NaN-preserving clamp)
            c0 = oq_code(w.x, q);
            c1 = oq_code(w.y, q);
          }
          const uint32_t packed = pack_bf2(oq_dequant(c0, q), oq_dequant(c1, q));
          *reinterpret_cast<uint32_t*>(dst + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4) = packed;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&b_full[rb.idx]);
          mbar_arrive(&w_empty[rw.idx]);
        }
        rw.next(SW);
        rb.next(SB);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile % p.num_m) * BM, n0 = (tile / p.num_m) * BN;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int row = m0 + h * 128 + q * 32 + lane;
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + h * 128);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          const int nc = n0 + c0;
          if (nc >= p.N) break;      // warp-uniform
          uint32_t r[32];
          tc_ld32(taddr + c0, r);
          tc_wait_ld();
          if (row < p.M) {
            const int valid = min(32, p.N - nc);
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < valid) v[j] += p.bias[nc + j];
            }
            if (p.res) {
              const bf16* rr = p.res + (long)row * p.ldr + nc;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < valid) v[j] += bf2f(rr[j]);
            }
            bf16* cr = p.C + (long)row * p.ldc + nc;
            if (valid == 32 && p.ldc % 8 == 0 && ((uintptr_t)p.C % 16) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                float t8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) t8[e] = v[j + e];
                store_bf8(cr + j, t8);
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
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// scale / zero point from saved group extrema and the LWC factors (quantize/quantizer.py:131-147)
__global__ void lwc_qparams_kernel(const float* xmin, const float* xmax, const float* up, const float* low, long n, int n_bits,
                                   int symmetric, int recip_div, float* scale, float* zp) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int lwc = up != nullptr;
  const OqQParams q = oq_qparams(xmin[i], xmax[i], lwc ? oq_sigmoidf(up[i]) : 1.0f, lwc ? oq_sigmoidf(low[i]) : 1.0f, lwc, n_bits,
                                 symmetric, recip_div);
  scale[i] = q.scale;
  zp[i] = q.zp;
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

}  // namespace

extern "C" int oq_lwc_qparams(const float* xmin, const float* xmax, const float* up, const float* low, long n, int n_bits,
                              int symmetric, int recip_div, float* scale_out, float* zp_out, cudaStream_t stream) {
  OQ_REQUIRE(xmin && xmax && scale_out && zp_out && n > 0, "oq_lwc_qparams: null pointer or empty");
  OQ_REQUIRE((up == nullptr) == (low == nullptr), "oq_lwc_qparams: up/low must both be set or both null");
  OQ_REQUIRE(n_bits >= 2 && n_bits <= 8, "oq_lwc_qparams: n_bits %d not in [2,8]", n_bits);
  lwc_qparams_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(xmin, xmax, up, low, n, n_bits, symmetric, recip_div, scale_out,
                                                                        zp_out);
  return oq_check_launch("oq_lwc_qparams");
}

extern "C" int oq_gemm_fq(const void* A, long lda, const float* W, long ldw, int M, int N, int K, const float* scale,
                          const float* zp, int group_size, int n_bits, void* C, long ldc, const float* bias,
                          const void* residual, long ldr, int max_ctas, cudaStream_t stream) {
  OQ_REQUIRE(A && W && C && scale && zp && M > 0 && N > 0 && K > 0, "oq_gemm_fq: null pointer or bad shape");
  OQ_REQUIRE(K % BK == 0, "oq_gemm_fq: K must be a multiple of %d", BK);
  OQ_REQUIRE(group_size == 0 || (group_size % BK == 0 && K % group_size == 0), "oq_gemm_fq: group_size must be 0 or a multiple of %d dividing K", BK);
  OQ_REQUIRE(n_bits >= 2 && n_bits <= 8, "oq_gemm_fq: n_bits %d not in [2,8]", n_bits);
  OQ_REQUIRE(lda % 8 == 0 && ldw % 4 == 0 && (uintptr_t)A % 16 == 0 && (uintptr_t)W % 16 == 0, "oq_gemm_fq: operands must be 16-byte aligned");
  EncodeTiledFn enc = get_encode();
  OQ_REQUIRE(enc, "cuTensorMapEncodeTiled not available from the driver");
  CUtensorMap ta, tw;
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)lda * 2};
    cuuint32_t box[2] = {64, 128}, estr[2] = {1, 1};
    CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(A), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    OQ_REQUIRE(r == CUDA_SUCCESS, "oq_gemm_fq: tensor map of A failed (%d)", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)ldw * 4};
    cuuint32_t box[2] = {64, 128}, estr[2] = {1, 1};
    CUresult r = enc(&tw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(W), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    OQ_REQUIRE(r == CUDA_SUCCESS, "oq_gemm_fq: tensor map of W failed (%d)", (int)r);
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_fq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) {
      oq_set_error("oq_gemm_fq: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr = true;
  }
  FqParams p = {};
  p.M = M; p.N = N; p.K = K; p.scale = scale; p.zp = zp; p.group = group_size; p.G = group_size ? K / group_size : 1;
  p.n_bits = n_bits; p.C = (bf16*)C; p.ldc = ldc; p.bias = bias; p.res = (const bf16*)residual; p.ldr = ldr;
  p.num_m = (M + BM - 1) / BM; p.num_n = (N + BN - 1) / BN;
  int grid = (int)std::min<long>((long)p.num_m * p.num_n, oq_num_sms());
  if (max_ctas > 0) grid = std::min(grid, max_ctas);
  gemm_fq_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tw, p);
  return oq_check_launch("oq_gemm_fq");
}
