// K5 / K6: fused causal attention of QuantMatMul -- QK^T, softmax and PV in one kernel, and a recompute backward -- on the
// 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA with 128-byte swizzle), sm_100a only.
//
// Reference: QuantLlamaAttention.forward, models/int_llama_layer.py:141-163 (qkt_matmul -> / sqrt(d) -> + mask ->
// fp32 softmax -> cast -> pv_matmul) and its autograd; OPT / Falcon call sites models/int_opt_layer.py:124-197,
// models/int_falcon_layer.py:148-176.  The operands are the already rotated + fake-quantised Q / K / V rows of the fused
// [tokens, q | k | v] matrix (csrc/act_quant.cu: aq_warp), so quantisation stays exactly where the reference has it.
//
// Why: the unfused chain (batched GEMM -> softmax kernel -> batched GEMM, and five launches in the backward) moves the
// [heads, T, T] scores / probabilities / their gradients through HBM: ~1.5 GB per optimiser step at LLaMA-7B shapes for
// ~0.2 GB of algorithmic Q / K / V / O traffic (profiles/r1_gemm_ncu_full_summary.csv).  Here nothing of size T x T
// leaves the SM: the forward keeps only two fp32 row statistics, the backward recomputes the probabilities tile by tile.
//
// Numerics follow the unfused path: scores are alpha * (q . k) rounded to fp16 (the reference's autocast dtype), the
// softmax is evaluated in fp32 with the FINAL row maximum and sum (two passes over the key tiles: P is normalised before
// it is rounded to bf16, like softmax-then-cast), P and dS are bf16 MMA operands, O / dQ / dK / dV are bf16.
//
//   forward   CTA = (sample, head, 128-query tile); key tiles j <= i.  pass 1: S_j -> m, l.  pass 2: S_j -> P_j -> O += P_j V_j.
//   dK / dV   CTA = (sample, K/V head, 128-key tile); loops over the group's query heads and the 64-query half tiles at or
//             below the diagonal; transposed formulation (S^T = K Q^T, dP^T = V dO^T) so that P^T and dS^T come out of TMEM
//             with one key row per thread and feed dV += P^T dO, dK += dS^T Q as K-major A operands.
//   dQ        CTA = (sample, head, 128-query tile); loops over 64-key half tiles; dQ += dS K.
// No atomics: every output element has one writer (dK / dV of a grouped-query K/V head are summed in TMEM).
//
// Warp roles (384 threads): warp 0 TMA producer, warp 1 MMA issuer (one thread), warp 2 TMEM allocator, warps 4-11 softmax /
// epilogue: TMEM lane = accumulator row = thread lane; the two warps of a lane quadrant (w and w + 4) split the columns.
// (One softmax warp per scheduler issued at IPC 0.38 -- fixed-latency and TMEM-load stalls with nothing to switch to,
// ncu profiles/r2_attn_ncu_summary.txt; two per scheduler hide most of it.)
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "mbar.cuh"

namespace {

constexpr int NUM_THREADS = 384;       // 4 control warps + 8 softmax warps
constexpr int SM_WARPS = 8;            // softmax / epilogue warps: two per TMEM lane quadrant, each takes half of the columns
constexpr int SM_THREADS = SM_WARPS * 32;
constexpr int ROW_B = 128;             // bytes per shared-memory row of a 64-column bf16 block (one swizzle row)
constexpr int BLK128 = 128 * ROW_B;    // 64-column block of a 128-row tile: 16 KB
constexpr int BLK64 = 64 * ROW_B;      // 64-column block of a  64-row tile:  8 KB
constexpr int PTILE = 2 * BLK128;      // a [128 x 128] bf16 probability tile (two 64-key blocks), whatever the head dimension
// Kernels are templates on the head dimension DH (64 or 128); inside them: NB = DH / 64 blocks per operand tile,
// KS = DH / 16 k-steps of a contraction over d, TILE128 / TILE64 = bytes of a [128 x DH] / [64 x DH] bf16 tile.
#define OQ_HEAD_DIM_CONSTS                                                                         \
  constexpr int D = DH, NB = DH / 64, KS = DH / 16, TILE128 = NB * BLK128, TILE64 = NB * BLK64;      \
  (void)NB; (void)KS; (void)TILE64; (void)TILE128

struct AttnParams {
  int T, heads, kv_heads, rep, batch, d;
  int cq, gq, ck, cv, gk;   // column of q head h: cq + (h / rep) * gq + (h % rep) * d;  k / v of K/V head g: ck / cv + g * gk
  float alpha;
  bf16* O; long ldo;                      // forward output [batch*T, heads*D]
  float* row_max; float* row_invl;        // [batch, heads, T]: max of the fp16 scores, 1 / sum exp(s - max)
  const float* delta;                     // [batch, heads, T]: rowsum(dO * O)
  bf16* dqkv; long ldg;                   // backward output, same column layout as the operand matrix
};

__device__ __forceinline__ int col_q(const AttnParams& p, int head) { return p.cq + (head / p.rep) * p.gq + (head % p.rep) * p.d; }
__device__ __forceinline__ int col_k(const AttnParams& p, int g) { return p.ck + g * p.gk; }
__device__ __forceinline__ int col_v(const AttnParams& p, int g) { return p.cv + g * p.gk; }

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// a [rows x DH] tile = NB blocks of 64 columns: one TMA box each
template <int NB>
__device__ __forceinline__ void load_tile(uint8_t* dst, int blk_bytes, const CUtensorMap* map, uint64_t* bar, int col, int row) {
#pragma unroll
  for (int b = 0; b < NB; ++b) tma_load_2d(dst + b * blk_bytes, map, bar, col + 64 * b, row);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// 8 columns (one 16-byte bf16 chunk of results per thread): the softmax warps read their accumulator slice in pieces so that the
// next piece is in flight while the current one is processed -- tcgen05.ld moves 64 B/clk per SM, so the 64 KB of an S / dP pair
// take ~1000 cycles, as long as the arithmetic on them; loading everything, waiting, then computing serialises the two
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// 64-bit shared-memory matrix descriptor, SWIZZLE_128B (same encoding as gemm_sm100.cu), kept as two 32-bit halves: the high
// word (SBO = 1024 B, version 1, 128-byte swizzle) is a constant, and every k-step / ring slot only ADDS a multiple of 16 bytes to
// the start-address field of the low word.  The MMA-issuing thread is the critical resource of these kernels (one thread issues
// 20-48 small MMAs per tile pair; with descriptors rebuilt by shift / mask / or chains and the TMEM address broadcast from a
// vector register every instruction cost ~150-200 cycles to issue -- profiles/r2_attn_ncu_summary.csv), so everything that
// is not `base + immediate` is hoisted out of the issue loops.
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
// low word of a K-major operand (rows of 128 B; LBO unused = 16 B)
__device__ __forceinline__ uint32_t klo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
// low word of an MN-major operand whose 64-column blocks are `blk` bytes apart (LBO)
__device__ __forceinline__ uint32_t mnlo(uint32_t saddr, uint32_t blk) { return ((saddr & 0x3FFFFu) >> 4) | ((blk >> 4) << 16); }
// address-field increment of k-step s (16 elements): K-major operand made of 64-column blocks `blk` bytes apart / MN-major (16 rows)
__device__ __forceinline__ constexpr uint32_t kstep(int s, uint32_t blk) { return ((uint32_t)(s >> 2) * blk + (uint32_t)(s & 3) * 32u) >> 4; }
__device__ __forceinline__ constexpr uint32_t mnstep(int s) { return (uint32_t)s * (16u * ROW_B >> 4); }
__device__ __forceinline__ void tc_mma2(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(DESC_HI), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor: D = f32, A = B = bf16, M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc(int n, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(b_mn ? 1 : 0) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// alpha * acc rounded to fp16 (the dtype the reference's autocast keeps the scores in), as fp32
__device__ __forceinline__ float score16(uint32_t bits, float alpha) {
  return __half2float(__float2half_rn(__uint_as_float(bits) * alpha));
}
// exp2 on the SFU without the denormal-range fix-up of exp2f / __expf (arguments are <= 0 here: results below 2^-126 flush to 0)
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float LOG2E = 1.4426950408889634f;
// The element loops below are written branch-free on purpose: `visible ? exp(..) : 0` compiles to a divergent branch with
// a reconvergence barrier PER ELEMENT (measured: 10x slower kernels); masking the SCORE with -inf (one select) and letting
// ex2(-inf) = 0 do the rest keeps one straight instruction stream, and tiles off the diagonal skip the select altogether.
template <bool MASK>
__device__ __forceinline__ float masked(float s, bool visible) { return (MASK && !visible) ? -INFINITY : s; }

// 8 values -> one 16-byte chunk of a swizzled 128-byte row (row r of a 1024-byte aligned block, logical chunk c)
__device__ __forceinline__ void st_chunk(uint8_t* blk, int r, int c, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf2(v[0], v[1]); u.y = pack_bf2(v[2], v[3]); u.z = pack_bf2(v[4], v[5]); u.w = pack_bf2(v[6], v[7]);
  *reinterpret_cast<uint4*>(blk + r * ROW_B + ((c ^ (r & 7)) << 4)) = u;
}

// Epilogue of a softmax warp: its [32 accumulator rows x COLS columns] slice from TMEM to global memory as bf16.  A thread owns
// one row (TMEM lane), so storing straight from registers makes every warp-wide 16-byte store touch 32 different rows (ncu on
// the dK / dV kernel: 17 % of the samples waited on the store queue).  The slice goes through shared memory (16-byte chunks
// XOR-swizzled by the row: conflict-free both ways) and leaves as whole rows: 32 lanes x 16 B = 256 / COLS rows per store.
// `stage`: 32 * COLS * 2 bytes of shared memory private to the warp and idle by now; gbase: row 0 / column 0 of the slice.
template <int COLS>
__device__ __forceinline__ void store_slice(uint4* stage, uint32_t taddr, float sc, bf16* gbase, long ld, int lane) {
  constexpr int CPR = COLS / 8, RPI = 32 / CPR;     // 16-byte chunks per row, rows per store instruction
#pragma unroll 1
  for (int c0 = 0; c0 < COLS; c0 += 32) {
    uint32_t v[32];
    tc_ld32(taddr + c0, v);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      float t8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) t8[e] = __uint_as_float(v[i + e]) * sc;
      const int ch = (c0 + i) >> 3;
      stage[lane * CPR + (ch ^ (lane & (CPR - 1) & 7))] =
          make_uint4(pack_bf2(t8[0], t8[1]), pack_bf2(t8[2], t8[3]), pack_bf2(t8[4], t8[5]), pack_bf2(t8[6], t8[7]));
    }
  }
  __syncwarp();
#pragma unroll 4
  for (int r0 = 0; r0 < 32; r0 += RPI) {
    const int r = r0 + lane / CPR, ch = lane % CPR;
    *reinterpret_cast<uint4*>(gbase + (long)r * ld + ch * 8) = stage[r * CPR + (ch ^ (r & (CPR - 1) & 7))];
  }
}

// ================================================================== forward
template <int DH>
constexpr int fwd_smem() { return 5 * (DH / 64) * BLK128 /*Q, K ring, V ring*/ + 2 * PTILE /*P x 2*/ + 1024 + 256; }

template <int DH>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm128, const AttnParams p) {
  OQ_HEAD_DIM_CONSTS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // K and V travel through SEPARATE two-slot rings: a K tile is released as soon as its S MMAs retire, a V tile only after
  // its P V MMAs -- one shared ring would hold K hostage until the softmax of its tile is done and expose the whole
  // TMA round trip in every iteration (measured: 7.2k cycles per tile pair instead of ~2.5k)
  uint8_t* sQ = smem;
  uint8_t* sKr = sQ + TILE128;                 // K ring: slot s at sKr + s * TILE128
  uint8_t* sVr = sKr + 2 * TILE128;            // V ring
  uint8_t* sP = sVr + 2 * TILE128;             // two P buffers: the softmax of tile j + 1 writes while P V of tile j reads
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * PTILE);
  uint64_t* q_full = bars;          // [1]
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* k_empty = bars + 3;     // [2]
  uint64_t* s_full = bars + 5;      // [2] MMA -> softmax
  uint64_t* s_empty = bars + 7;     // [2] softmax -> MMA (one arrive per softmax warp)
  uint64_t* p_full = bars + 9;      // [2] softmax -> MMA
  uint64_t* p_empty = bars + 11;    // [2] MMA -> softmax
  uint64_t* o_full = bars + 13;     // [1]
  uint64_t* v_full = bars + 14;     // [2]
  uint64_t* v_empty = bars + 16;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = p.T / 128, per = p.heads * p.batch;
  const int qt = nt - 1 - (int)blockIdx.x / per;          // heavy (late) query tiles first
  const int rem = (int)blockIdx.x % per;
  const int head = rem % p.heads, b = rem / p.heads;
  const int g = head / p.rep;
  const int row0 = b * p.T;
  const int nj = qt + 1, nit = 2 * nj;
  const int qc = col_q(p, head), kc = col_k(p, g), vc = col_v(p, g);
  // shared-memory base as a value the compiler KNOWS to be warp-uniform (lane-0 broadcast): the descriptors derived from it
  // stay in uniform registers and the MMA issue loop needs no per-instruction register -> uniform-register transfer
  const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);

  if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm128) : "memory");
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], SM_WARPS);
      mbar_init(&p_full[i], SM_WARPS);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // The CTA owns all 512 TMEM columns, so the allocation starts at column 0 / lane 0: TMEM addresses are literals below
  // (a register-held base would have to be broadcast into a uniform register before every MMA).
  if (*tmem_slot != 0u) __trap();
  constexpr uint32_t tmem = 0;           // S0 [0,128)  S1 [128,256)  O [256,384)

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, TILE128);
      load_tile<NB>(sQ, BLK128, &tm128, q_full, qc, row0 + qt * 128);
      for (int t = 0; t < nit; ++t) {
        const int st = t & 1, pass = t >= nj, j = pass ? t - nj : t;
        mbar_wait(&k_empty[st], ((t >> 1) & 1) ^ 1);
        uint8_t* sK = sKr + st * TILE128;
        mbar_expect_tx(&k_full[st], TILE128);
        load_tile<NB>(sK, BLK128, &tm128, &k_full[st], kc, row0 + j * 128);
        if (pass) {
          const int vs = j & 1;
          mbar_wait(&v_empty[vs], ((j >> 1) & 1) ^ 1);
          uint8_t* sV = sVr + vs * TILE128;
          mbar_expect_tx(&v_full[vs], TILE128);
          load_tile<NB>(sV, BLK128, &tm128, &v_full[vs], vc, row0 + j * 128);
        }
      }
    }
  } else if (warp == 1) {
    // two MMA-issuing threads (a tcgen05.mma costs its issuing thread ~100 cycles, see attn_dq_kernel): S = Q K^T of both
    // passes here, O += P V in warp 3
    if (lane == 0) {
      const uint32_t id_s = make_idesc(128, 0);
      const uint32_t qlo = klo(sbase), klo0 = klo(sbase + TILE128);
      mbar_wait(q_full, 0);
      for (int t = 0; t < nit; ++t) {
        const int st = t & 1;
        const uint32_t ph = (t >> 1) & 1;
        mbar_wait(&k_full[st], ph);
        mbar_wait(&s_empty[st], ph ^ 1);
        tc_fence_after();
        const uint32_t kl = klo0 + st * (TILE128 >> 4);
        tc_mma2(st * 128, qlo, kl, id_s, 0u);
#pragma unroll
        for (int s = 1; s < KS; ++s) tc_mma2(st * 128, qlo + kstep(s, BLK128), kl + kstep(s, BLK128), id_s, 1u);
        tc_commit(&k_empty[st]);
        tc_commit(&s_full[st]);
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {
      const uint32_t id_o = make_idesc(D, 1);
      const uint32_t plo = klo(sbase + 5 * TILE128), vlo0 = mnlo(sbase + 3 * TILE128, BLK128);
      for (int j = 0; j < nj; ++j) {        // O += P(j) V(j), key tile j of pass 2
        const int vs = j & 1;
        const uint32_t vlo = vlo0 + vs * (TILE128 >> 4), pl = plo + vs * (PTILE >> 4);
        mbar_wait(&v_full[vs], (j >> 1) & 1);
        mbar_wait(&p_full[vs], (j >> 1) & 1);
        tc_fence_after();
        tc_mma2(256, pl, vlo, id_o, j > 0 ? 1u : 0u);
#pragma unroll
        for (int s = 1; s < 8; ++s) tc_mma2(256, pl + kstep(s, BLK128), vlo + mnstep(s), id_o, 1u);
        tc_commit(&p_empty[vs]);
        tc_commit(&v_empty[vs]);
      }
      tc_commit(o_full);
    }
  } else if (warp >= 4) {
    const int q = warp & 3, r = q * 32 + lane, half = (warp - 4) >> 2;     // this warp's columns: [64 half, 64 half + 64)
    const int qrow = qt * 128 + r;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    float m = -INFINITY, l = 0.f;      // row maximum of the fp16 scores and sum exp(s - m) over this warp's columns
    // ---- pass 1: row maximum and sum
    auto stats_tile = [&](int sb, auto mask_tag) {
      constexpr bool MASK = decltype(mask_tag)::value;
#pragma unroll 1
      for (int c0 = 64 * half; c0 < 64 * half + 64; c0 += 32) {
        uint32_t v[32];
        tc_ld32(lane_base + sb * 128 + c0, v);
        tc_wait_ld();
        float sc[32], mx = m;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          sc[i] = masked<MASK>(score16(v[i], p.alpha), c0 + i <= r);     // causal: keys c <= r on the diagonal tile
          mx = fmaxf(mx, sc[i]);
        }
        // a warp whose columns are all masked so far (second column half, first rows of the first query tile) keeps
        // mx = -inf: use 0 as the reference then (every term is exp2(-inf) = 0) instead of forming inf - inf
        const float mx2 = (mx == -INFINITY) ? 0.f : mx * LOG2E;
        float add0 = 0.f, add1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          add0 += ex2_ftz(fmaf(sc[i], LOG2E, -mx2));
          add1 += ex2_ftz(fmaf(sc[i + 1], LOG2E, -mx2));
        }
        l = l * ((m == mx) ? 1.0f : ex2_ftz((m - mx) * LOG2E)) + (add0 + add1);   // m = -inf < mx: l = 0 * 0
        m = mx;
      }
    };
    for (int t = 0; t < nj; ++t) {
      const int sb = t & 1;
      mbar_wait(&s_full[sb], (t >> 1) & 1);
      tc_fence_after();
      if (t == qt) stats_tile(sb, std::true_type{});
      else stats_tile(sb, std::false_type{});
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);
    }
    {   // merge the two column halves of every row: (m, l) pairs meet through shared memory (sP is still unused)
      float2* xch = reinterpret_cast<float2*>(sP);
      xch[half * 128 + r] = make_float2(m, l);
      named_bar_sync(1, SM_THREADS);
      const float2 o = xch[(half ^ 1) * 128 + r];
      const float mf = fmaxf(m, o.x);       // finite: key 0 is visible to every row and lies in half 0
      l = l * ex2_ftz((m - mf) * LOG2E) + o.y * ex2_ftz((o.x - mf) * LOG2E);
      m = mf;
      named_bar_sync(1, SM_THREADS);        // everybody has read the exchange area before pass 2 writes P over it
    }
    const float inv = 1.0f / l;
    const long srow = ((long)b * p.heads + head) * p.T + qrow;
    if (half == 0) {
      p.row_max[srow] = m;
      p.row_invl[srow] = inv;
    }
    const float m2 = m * LOG2E;
    // ---- pass 2: probabilities with the final statistics -> shared memory (A operand of P V)
    auto prob_tile = [&](int sb, uint8_t* pbuf, auto mask_tag) {
      constexpr bool MASK = decltype(mask_tag)::value;
#pragma unroll 1
      for (int c0 = 64 * half; c0 < 64 * half + 64; c0 += 32) {
        uint32_t v[32];
        tc_ld32(lane_base + sb * 128 + c0, v);
        tc_wait_ld();
        uint8_t* blk = pbuf + (c0 >> 6) * BLK128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float pr[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = c0 + ch * 8 + e;
            pr[e] = ex2_ftz(fmaf(masked<MASK>(score16(v[ch * 8 + e], p.alpha), c <= r), LOG2E, -m2)) * inv;
          }
          st_chunk(blk, r, ((c0 & 63) >> 3) + ch, pr);
        }
      }
    };
    for (int t = nj; t < nit; ++t) {
      const int sb = t & 1, j = t - nj;
      mbar_wait(&s_full[sb], (t >> 1) & 1);
      tc_fence_after();
      const int pb = j & 1;
      mbar_wait(&p_empty[pb], ((j >> 1) & 1) ^ 1);      // P V of tile j - 2 has finished reading this P buffer
      if (j == qt) prob_tile(sb, sP + pb * PTILE, std::true_type{});
      else prob_tile(sb, sP + pb * PTILE, std::false_type{});
      fence_proxy_async();      // generic-proxy writes of sP -> visible to the tensor core's async proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&p_full[pb]);
        mbar_arrive(&s_empty[sb]);
      }
    }
    // ---- epilogue
    mbar_wait(o_full, 0);
    tc_fence_after();
    // this warp: rows [32 q, +32) of the query tile, columns [D/2 half, +D/2) of the head; the Q tile is idle by now
    store_slice<D / 2>(reinterpret_cast<uint4*>(sQ) + (warp - 4) * (32 * (D / 16)), lane_base + 256 + (D / 2) * half, 1.0f,
                       p.O + (long)(row0 + qt * 128 + q * 32) * p.ldo + head * D + (D / 2) * half, p.ldo, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// ================================================================== backward: delta = rowsum(dO * O)
template <int DH>
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* O, long ldo, const bf16* dO, long lddo, float* delta, int batch,
                                                         int T, int heads) {
  constexpr int D = DH;
  const long w = (long)blockIdx.x * 8 + (threadIdx.x >> 5);      // one warp per (token row, head)
  const int lane = threadIdx.x & 31;
  const long total = (long)batch * T * heads;
  if (w >= total) return;
  const long row = w / heads;
  const int head = (int)(w % heads);
  float s;
  if constexpr (D == 128) {
    float a[4], c[4];
    load_bf4(O + row * ldo + head * D + lane * 4, a);
    load_bf4(dO + row * lddo + head * D + lane * 4, c);
    s = a[0] * c[0] + a[1] * c[1] + a[2] * c[2] + a[3] * c[3];
  } else {
    float a0, a1, c0, c1;
    unpack_bf2(*reinterpret_cast<const uint32_t*>(O + row * ldo + head * D + lane * 2), a0, a1);
    unpack_bf2(*reinterpret_cast<const uint32_t*>(dO + row * lddo + head * D + lane * 2), c0, c1);
    s = a0 * c0 + a1 * c1;
  }
  s = warp_reduce(s, OpSum());
  if (lane == 0) {
    const long bi = row / T, t = row % T;
    delta[(bi * heads + head) * T + t] = s;
  }
}

// ================================================================== backward: dK, dV
constexpr int DKV_ST = 3;      // Q / dO stages: a stage is needed twice (S^T / dP^T early, the accumulate MMAs after the softmax),
                               // so it is held for a whole iteration; three stages keep the TMA round trip off the critical path
template <int DH>
constexpr int dkv_smem() {
  return 2 * (DH / 64) * BLK128 /*K,V*/ + DKV_ST * 2 * (DH / 64) * BLK64 /*Q,dO*/ + 2 * BLK128 /*P^T,dS^T*/ + SM_WARPS * 2 * 32 * 16 /*stats*/ + 1024 + 256;
}

template <int DH>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_dkv_kernel(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tm64,
                const __grid_constant__ CUtensorMap tmdo64, const AttnParams p) {
  OQ_HEAD_DIM_CONSTS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sK = smem;
  uint8_t* sV = sK + TILE128;
  uint8_t* sQdO = sV + TILE128;               // stage s: Q at + s * 2 * TILE64, dO right behind it
  uint8_t* sPS = sQdO + DKV_ST * 2 * TILE64;  // P^T, dS^T right behind it (one buffer: the accumulate MMAs read it for ~500 cycles)
  float4* sstat = reinterpret_cast<float4*>(sPS + 2 * BLK128);   // [softmax warp][2][32]: (max*log2e, 1/l, delta, -) of the warp's 32 query columns
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sstat) + SM_WARPS * 2 * 32 * 16);
  uint64_t* kv_full = bars;          // [1]
  uint64_t* qdo_full = bars + 1;     // [DKV_ST]
  uint64_t* qdo_empty = bars + 4;    // [DKV_ST]
  uint64_t* s_full = bars + 7;       // [2]
  uint64_t* s_empty = bars + 9;      // [2]
  uint64_t* p_full = bars + 11;      // [1]
  uint64_t* p_empty = bars + 12;     // [1]
  uint64_t* acc_full = bars + 13;    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = p.T / 128, per = p.kv_heads * p.batch;
  const int j = (int)blockIdx.x / per;                      // key tile; early tiles see the most queries: heavy first
  const int rem = (int)blockIdx.x % per;
  const int g = rem % p.kv_heads, b = rem / p.kv_heads;
  const int row0 = b * p.T;
  const int nh = 2 * nt - 2 * j;                            // 64-query half tiles at or below the diagonal
  const int nit = p.rep * nh;
  const int kc = col_k(p, g), vc = col_v(p, g);
  const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);     // warp-uniform by construction (see attn_fwd_kernel)

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm128) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm64) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmdo64) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < DKV_ST; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 2);      // one commit per issuer (S^T; dP^T)
      mbar_init(&s_empty[i], SM_WARPS);
    }
    mbar_init(p_full, SM_WARPS);
    mbar_init(p_empty, 1);
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (*tmem_slot != 0u) __trap();      // whole TMEM allocated: base 0 (see attn_fwd_kernel)
  constexpr uint32_t tmem = 0;         // buffer s: S^T [s*128, +64)  dP^T [s*128+64, +64);  dK [256,384)  dV [384,512)

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * TILE128);
      load_tile<NB>(sK, BLK128, &tm128, kv_full, kc, row0 + j * 128);
      load_tile<NB>(sV, BLK128, &tm128, kv_full, vc, row0 + j * 128);
      for (int it = 0; it < nit; ++it) {
        const int st = it % DKV_ST, head = g * p.rep + it / nh, i64 = 2 * j + it % nh;
        mbar_wait(&qdo_empty[st], ((it / DKV_ST) & 1) ^ 1);
        uint8_t* sQ = sQdO + st * 2 * TILE64;
        mbar_expect_tx(&qdo_full[st], 2 * TILE64);
        const int qc = col_q(p, head), r = row0 + i64 * 64;
        load_tile<NB>(sQ, BLK64, &tm64, &qdo_full[st], qc, r);
        load_tile<NB>(sQ + TILE64, BLK64, &tmdo64, &qdo_full[st], head * D, r);
      }
    }
  } else if (warp == 1) {
    // three MMA-issuing threads (see attn_dq_kernel): S^T = K Q^T here, dP^T = V dO^T in warp 3, the accumulations in warp 2
    if (lane == 0) {
      const uint32_t id_s = make_idesc(64, 0);
      const uint32_t kl = klo(sbase), q_k0 = klo(sbase + 2 * TILE128);     // stage 0: Q, then dO at + TILE64
      mbar_wait(kv_full, 0);
      for (int it = 0; it < nit; ++it) {
        const int st = it % DKV_ST, sb = it & 1;
        mbar_wait(&qdo_full[st], (it / DKV_ST) & 1);
        mbar_wait(&s_empty[sb], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t qk = q_k0 + st * (2 * TILE64 >> 4);
        tc_mma2(sb * 128, kl, qk, id_s, 0u);
#pragma unroll
        for (int s = 1; s < KS; ++s) tc_mma2(sb * 128, kl + kstep(s, BLK128), qk + kstep(s, BLK64), id_s, 1u);
        tc_commit(&s_full[sb]);
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {
      const uint32_t id_s = make_idesc(64, 0);
      const uint32_t vl = klo(sbase + TILE128), q_k0 = klo(sbase + 2 * TILE128);
      mbar_wait(kv_full, 0);
      for (int it = 0; it < nit; ++it) {
        const int st = it % DKV_ST, sb = it & 1;
        mbar_wait(&qdo_full[st], (it / DKV_ST) & 1);
        mbar_wait(&s_empty[sb], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t dok = q_k0 + st * (2 * TILE64 >> 4) + (TILE64 >> 4);
        tc_mma2(sb * 128 + 64, vl, dok, id_s, 0u);
#pragma unroll
        for (int s = 1; s < KS; ++s) tc_mma2(sb * 128 + 64, vl + kstep(s, BLK128), dok + kstep(s, BLK64), id_s, 1u);
        tc_commit(&s_full[sb]);
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {       // dV += P^T dO,  dK += dS^T Q   (K = the 64 queries of stage u)
      const uint32_t id_a = make_idesc(D, 1);
      const uint32_t q_mn0 = mnlo(sbase + 2 * TILE128, BLK64);
      const uint32_t ptlo = klo(sbase + 2 * TILE128 + DKV_ST * 2 * TILE64), dstlo = ptlo + (BLK128 >> 4);
      for (int u = 0; u < nit; ++u) {
        const int st = u % DKV_ST;
        const uint32_t qmn = q_mn0 + st * (2 * TILE64 >> 4), domn = qmn + (TILE64 >> 4);
        const uint32_t first = u > 0 ? 1u : 0u;
        mbar_wait(&qdo_full[st], (u / DKV_ST) & 1);
        mbar_wait(p_full, u & 1);
        tc_fence_after();
        tc_mma2(256 + D, ptlo, domn, id_a, first);
#pragma unroll
        for (int s = 1; s < 4; ++s) tc_mma2(256 + D, ptlo + kstep(s, BLK128), domn + mnstep(s), id_a, 1u);
        tc_mma2(256, dstlo, qmn, id_a, first);
#pragma unroll
        for (int s = 1; s < 4; ++s) tc_mma2(256, dstlo + kstep(s, BLK128), qmn + mnstep(s), id_a, 1u);
        tc_commit(p_empty);
        tc_commit(&qdo_empty[st]);      // the S^T / dP^T MMAs on this stage (other issuers) retired before P^T existed
      }
      tc_commit(acc_full);
    }
  } else if (warp >= 4) {
    const int q = warp & 3, kr = q * 32 + lane, half = (warp - 4) >> 2;   // key row of this thread; query columns [32 half, +32)
    const int kg = j * 128 + kr;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    // per-query statistics (max * log2e, 1/l, delta) of the 32 query columns this warp works on: every warp keeps its own
    // two-slot copy in shared memory (lane l fetches column 32 half + l), loaded ONE ITERATION AHEAD into registers and
    // stored after the tile work, so neither the global-load latency nor a CTA-wide barrier sits between two tiles (ncu on the
    // version that staged them once per CTA behind a named barrier: barrier 2.8 + long scoreboard 2.1 stalled warps per issue)
    float4* wstat = sstat + (warp - 4) * 64;
    auto stat_row = [&](int it2) {
      const int head2 = g * p.rep + it2 / nh, i64b = 2 * j + it2 % nh;
      return ((long)b * p.heads + head2) * p.T + i64b * 64 + 32 * half + lane;
    };
    float nx0 = 0.f, nx1 = 0.f, nx2 = 0.f;
    if (nit > 0) {
      const long sr = stat_row(0);
      wstat[lane] = make_float4(p.row_max[sr] * LOG2E, p.row_invl[sr], p.delta[sr], 0.f);
    }
    __syncwarp();
    for (int it = 0; it < nit; ++it) {
      const int sb = it & 1, ss = it & 1, i64 = 2 * j + it % nh;
      const uint32_t ph = (it >> 1) & 1;
      const bool pre = it + 1 < nit;
      if (pre) {
        const long sr = stat_row(it + 1);
        nx0 = p.row_max[sr]; nx1 = p.row_invl[sr]; nx2 = p.delta[sr];
      }
      mbar_wait(&s_full[sb], ph);
      tc_fence_after();
      uint8_t* bPT = sPS;
      uint8_t* bdST = bPT + BLK128;
      const int lo = kg - i64 * 64;            // causal: query column c is visible to this key row iff c >= lo
      auto tile = [&](auto mask_tag) {
        constexpr bool MASK = decltype(mask_tag)::value;
        const int c0 = 32 * half;
        uint32_t sv[32], dv[32];
        tc_ld32(lane_base + sb * 128 + c0, sv);
        tc_ld32(lane_base + sb * 128 + 64 + c0, dv);
        tc_wait_ld();
        uint4 pk[4], dk[4];       // packed bf16 results: computed while the previous step's accumulate MMAs still read the buffer
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float pr[8], ds[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = c0 + ch * 8 + e;
            const float4 stt = wstat[ss * 32 + ch * 8 + e];   // (max * log2e, 1/l, delta, -) of query column c
            const float pe = ex2_ftz(fmaf(masked<MASK>(score16(sv[ch * 8 + e], p.alpha), c >= lo), LOG2E, -stt.x)) * stt.y;
            pr[e] = pe;
            ds[e] = pe * (__uint_as_float(dv[ch * 8 + e]) - stt.z);
          }
          pk[ch] = make_uint4(pack_bf2(pr[0], pr[1]), pack_bf2(pr[2], pr[3]), pack_bf2(pr[4], pr[5]), pack_bf2(pr[6], pr[7]));
          dk[ch] = make_uint4(pack_bf2(ds[0], ds[1]), pack_bf2(ds[2], ds[3]), pack_bf2(ds[4], ds[5]), pack_bf2(ds[6], ds[7]));
        }
        if (it > 0) mbar_wait(p_empty, (it - 1) & 1);     // the accumulate MMAs of the previous step have finished reading P^T / dS^T
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int off = kr * ROW_B + ((((c0 >> 3) + ch) ^ (kr & 7)) << 4);
          *reinterpret_cast<uint4*>(bPT + off) = pk[ch];
          *reinterpret_cast<uint4*>(bdST + off) = dk[ch];
        }
      };
      if (i64 < 2 * j + 2) tile(std::true_type{});      // the two half tiles that touch the diagonal
      else tile(std::false_type{});
      // the other slot was last read in iteration it - 1, which every lane left through the __syncwarp below
      if (pre) wstat[((it + 1) & 1) * 32 + lane] = make_float4(nx0 * LOG2E, nx1, nx2, 0.f);
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(p_full);
        mbar_arrive(&s_empty[sb]);
      }
    }
    // ---- epilogue: dK (x alpha: s = alpha q.k) by the warps of column half 0, dV by half 1; the Q / dO ring is idle by now
    mbar_wait(acc_full, 0);
    tc_fence_after();
    store_slice<D>(reinterpret_cast<uint4*>(sQdO) + (warp - 4) * (32 * (D / 8)), lane_base + 256 + D * half, half == 0 ? p.alpha : 1.0f,
                   p.dqkv + (long)(row0 + j * 128 + q * 32) * p.ldg + (half == 0 ? kc : vc), p.ldg, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// ================================================================== backward: dQ
constexpr int DQ_KST = 4;      // K ring: a K half tile serves S early and dQ += dS K after the softmax (held for a whole iteration)
template <int DH>
constexpr int dq_smem() {
  return 2 * (DH / 64) * BLK128 /*Q,dO*/ + (DQ_KST + 2) * (DH / 64) * BLK64 /*K ring, V ring*/ + 2 * BLK128 /*dS x 2*/ + 1024 + 256;
}

template <int DH>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_dq_kernel(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tm64,
               const __grid_constant__ CUtensorMap tmdo128, const AttnParams p) {
  OQ_HEAD_DIM_CONSTS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + TILE128;
  uint8_t* sKr = sdO + TILE128;                // K ring: slot s at + s * TILE64
  uint8_t* sVr = sKr + DQ_KST * TILE64;        // V ring (2 slots): V only feeds dP and is released with the S / dP MMAs
  uint8_t* sdS = sVr + 2 * TILE64;             // buffer s at + s * BLK128
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + 2 * BLK128);
  uint64_t* qdo_full = bars;         // [1]
  uint64_t* k_full = bars + 1;       // [DQ_KST]
  uint64_t* k_empty = bars + 5;      // [DQ_KST]
  uint64_t* v_full = bars + 9;       // [2]
  uint64_t* v_empty = bars + 11;     // [2]
  uint64_t* s_full = bars + 13;      // [2]
  uint64_t* s_empty = bars + 15;     // [2]
  uint64_t* p_full = bars + 17;      // [2]
  uint64_t* p_empty = bars + 19;     // [2]
  uint64_t* acc_full = bars + 21;    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = p.T / 128, per = p.heads * p.batch;
  const int i = nt - 1 - (int)blockIdx.x / per;            // heavy (late) query tiles first
  const int rem = (int)blockIdx.x % per;
  const int head = rem % p.heads, b = rem / p.heads;
  const int g = head / p.rep;
  const int row0 = b * p.T;
  const int nit = 2 * i + 2;                               // 64-key half tiles up to the diagonal
  const int qc = col_q(p, head), kc = col_k(p, g), vc = col_v(p, g);
  const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);     // warp-uniform by construction (see attn_fwd_kernel)

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm128) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm64) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmdo128) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(qdo_full, 1);
    for (int k = 0; k < DQ_KST; ++k) {
      mbar_init(&k_full[k], 1);
      mbar_init(&k_empty[k], 1);
    }
    for (int k = 0; k < 2; ++k) {
      mbar_init(&v_full[k], 1);
      mbar_init(&v_empty[k], 1);
      mbar_init(&s_full[k], 2);      // one commit per MMA-issuing thread (S; dP)
      mbar_init(&s_empty[k], SM_WARPS);
      mbar_init(&p_full[k], SM_WARPS);
      mbar_init(&p_empty[k], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (*tmem_slot != 0u) __trap();      // whole TMEM allocated: base 0 (see attn_fwd_kernel)
  constexpr uint32_t tmem = 0;         // buffer s: S [s*128, +64)  dP [s*128+64, +64);  dQ [256,384)

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(qdo_full, 2 * TILE128);
      load_tile<NB>(sQ, BLK128, &tm128, qdo_full, qc, row0 + i * 128);
      load_tile<NB>(sdO, BLK128, &tmdo128, qdo_full, head * D, row0 + i * 128);
      for (int it = 0; it < nit; ++it) {
        const int ks = it % DQ_KST, vs = it & 1;
        const int r = row0 + it * 64;
        mbar_wait(&k_empty[ks], ((it / DQ_KST) & 1) ^ 1);
        uint8_t* sK = sKr + ks * TILE64;
        mbar_expect_tx(&k_full[ks], TILE64);
        load_tile<NB>(sK, BLK64, &tm64, &k_full[ks], kc, r);
        mbar_wait(&v_empty[vs], ((it >> 1) & 1) ^ 1);
        uint8_t* sV = sVr + vs * TILE64;
        mbar_expect_tx(&v_full[vs], TILE64);
        load_tile<NB>(sV, BLK64, &tm64, &v_full[vs], vc, r);
      }
    }
  } else if (warp == 1) {
    // Two MMA-issuing threads.  A tcgen05.mma of these small shapes (N = 64: 32 cycles of tensor work) costs the issuing
    // thread ~100+ cycles, so ONE issuer bounded the kernel (ncu: the softmax warps spent 25 % of their samples waiting for
    // S, the tensor pipe was 30 % active): this warp issues S = Q K^T, warp 3 dP = dO V^T, warp 2 dQ += dS K.
    if (lane == 0) {
      const uint32_t id_s = make_idesc(64, 0);
      const uint32_t ql = klo(sbase), k_k0 = klo(sbase + 2 * TILE128);
      mbar_wait(qdo_full, 0);
      for (int it = 0; it < nit; ++it) {
        const int st = it & 1, ks = it % DQ_KST;
        mbar_wait(&k_full[ks], (it / DQ_KST) & 1);
        mbar_wait(&s_empty[st], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t kk = k_k0 + ks * (TILE64 >> 4);
        tc_mma2(st * 128, ql, kk, id_s, 0u);
#pragma unroll
        for (int s = 1; s < KS; ++s) tc_mma2(st * 128, ql + kstep(s, BLK128), kk + kstep(s, BLK64), id_s, 1u);
        tc_commit(&s_full[st]);
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {       // dP = dO V^T
      const uint32_t id_s = make_idesc(64, 0);
      const uint32_t dol = klo(sbase + TILE128), v_k0 = klo(sbase + 2 * TILE128 + DQ_KST * TILE64);
      mbar_wait(qdo_full, 0);
      for (int it = 0; it < nit; ++it) {
        const int st = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(&v_full[st], ph);
        mbar_wait(&s_empty[st], ph ^ 1);
        tc_fence_after();
        const uint32_t vk = v_k0 + st * (TILE64 >> 4);
        tc_mma2(st * 128 + 64, dol, vk, id_s, 0u);
#pragma unroll
        for (int s = 1; s < KS; ++s) tc_mma2(st * 128 + 64, dol + kstep(s, BLK128), vk + kstep(s, BLK64), id_s, 1u);
        tc_commit(&v_empty[st]);
        tc_commit(&s_full[st]);
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {       // dQ += dS K   (K = the 64 keys of step u); third issuer: the TMEM-allocating warp is idle in the main loop
      const uint32_t id_a = make_idesc(D, 1);
      const uint32_t k_mn0 = mnlo(sbase + 2 * TILE128, BLK64), ds0 = klo(sbase + 2 * TILE128 + (DQ_KST + 2) * TILE64);
      for (int u = 0; u < nit; ++u) {
        const int sb = u & 1, ks = u % DQ_KST;
        const uint32_t kmn = k_mn0 + ks * (TILE64 >> 4), dsl = ds0 + sb * (BLK128 >> 4);
        mbar_wait(&k_full[ks], (u / DQ_KST) & 1);          // (the S issuer waited for it too; this thread reads K as well)
        mbar_wait(&p_full[sb], (u >> 1) & 1);
        tc_fence_after();
        tc_mma2(256, dsl, kmn, id_a, u > 0 ? 1u : 0u);
#pragma unroll
        for (int s = 1; s < 4; ++s) tc_mma2(256, dsl + kstep(s, BLK128), kmn + mnstep(s), id_a, 1u);
        tc_commit(&p_empty[sb]);
        tc_commit(&k_empty[ks]);     // the S MMAs on this K slot (other issuer) retired before dS existed
      }
      tc_commit(acc_full);
    }
  } else if (warp >= 4) {
    const int q = warp & 3, r = q * 32 + lane, half = (warp - 4) >> 2;     // key columns [32 half, +32) of every half tile
    const int qg = i * 128 + r;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const long sr = ((long)b * p.heads + head) * p.T + qg;
    const float m2 = p.row_max[sr] * LOG2E, inv = p.row_invl[sr], dl = p.delta[sr];
    for (int it = 0; it < nit; ++it) {
      const int sb = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(&s_full[sb], ph);
      tc_fence_after();
      mbar_wait(&p_empty[sb], ph ^ 1);
      uint8_t* bdS = sdS + sb * BLK128;
      const int hi = qg - it * 64;             // causal: key column c is visible to this query row iff c <= hi
      auto tile = [&](auto mask_tag) {
        constexpr bool MASK = decltype(mask_tag)::value;
        {
          const int c0 = 32 * half;
          uint32_t sv[2][8], dv[2][8];
          tc_ld8(lane_base + sb * 128 + c0, sv[0]);
          tc_ld8(lane_base + sb * 128 + 64 + c0, dv[0]);
          tc_wait_ld();
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            if (ch < 3) {      // next 8 columns in flight while these are processed
              tc_ld8(lane_base + sb * 128 + c0 + 8 * (ch + 1), sv[(ch + 1) & 1]);
              tc_ld8(lane_base + sb * 128 + 64 + c0 + 8 * (ch + 1), dv[(ch + 1) & 1]);
            }
            float ds[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int c = c0 + ch * 8 + e;
              const float pe = ex2_ftz(fmaf(masked<MASK>(score16(sv[ch & 1][e], p.alpha), c <= hi), LOG2E, -m2)) * inv;
              ds[e] = pe * (__uint_as_float(dv[ch & 1][e]) - dl);
            }
            st_chunk(bdS, r, (c0 >> 3) + ch, ds);
            if (ch < 3) tc_wait_ld();
          }
        }
      };
      if (it >= 2 * i) tile(std::true_type{});           // the two half tiles that touch the diagonal
      else tile(std::false_type{});
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&p_full[sb]);
        mbar_arrive(&s_empty[sb]);
      }
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    // dQ (x alpha): rows [32 q, +32) of the query tile, columns [D/2 half, +D/2); the Q tile is idle by now
    store_slice<D / 2>(reinterpret_cast<uint4*>(sQ) + (warp - 4) * (32 * (D / 16)), lane_base + 256 + (D / 2) * half, p.alpha,
                       p.dqkv + (long)(row0 + i * 128 + q * 32) * p.ldg + qc + (D / 2) * half, p.ldg, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
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
// 2-D bf16 map over a [rows, cols] row-major matrix (row stride ld elements); box 64 columns x box_rows rows
int make_map(CUtensorMap* map, const void* base, long cols, long rows, long ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  OQ_REQUIRE(enc, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  OQ_REQUIRE(r == CUDA_SUCCESS, "attention: cuTensorMapEncodeTiled failed (%d): base=%p cols=%ld rows=%ld ld=%ld", (int)r, base, cols, rows, ld);
  return 0;
}
template <typename K>
int set_smem(K kernel, int bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    oq_set_error("%s: cudaFuncSetAttribute failed: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}
int check_common(const char* fn, const void* qkv, long ld, int batch, int T, int heads, int kv_heads, int head_dim, int cq, int gq,
                 int ck, int cv, int gk, long width) {
  OQ_REQUIRE(qkv && batch > 0 && heads > 0 && kv_heads > 0 && heads % kv_heads == 0, "%s: null pointer or bad head counts", fn);
  OQ_REQUIRE((head_dim == 64 || head_dim == 128) && T >= 128 && T % 128 == 0, "%s: needs head_dim 64 or 128 and a sequence length that "
             "is a multiple of 128 (got d=%d T=%d); use the unfused GEMM + softmax path otherwise", fn, head_dim, T);
  const int D = head_dim;
  OQ_REQUIRE(ld % 8 == 0 && (uintptr_t)qkv % 16 == 0 && cq % 8 == 0 && gq % 8 == 0 && ck % 8 == 0 && cv % 8 == 0 && gk % 8 == 0,
             "%s: operand matrix must be 16-byte aligned with column offsets / strides that are multiples of 8", fn);
  const int rep = heads / kv_heads;
  const long last_q = cq + (long)(kv_heads - 1) * gq + (long)(rep - 1) * D + D;
  const long last_k = ck + (long)(kv_heads - 1) * gk + D, last_v = cv + (long)(kv_heads - 1) * gk + D;
  OQ_REQUIRE(last_q <= width && last_k <= width && last_v <= width && width <= ld, "%s: head columns exceed the matrix width", fn);
  return 0;
}
AttnParams base_params(int batch, int T, int heads, int kv_heads, int d, int cq, int gq, int ck, int cv, int gk, float alpha) {
  AttnParams p = {};
  p.T = T; p.heads = heads; p.kv_heads = kv_heads; p.rep = heads / kv_heads; p.batch = batch; p.d = d;
  p.cq = cq; p.gq = gq; p.ck = ck; p.cv = cv; p.gk = gk; p.alpha = alpha;
  return p;
}

}  // namespace

extern "C" int oq_attn_supported(int T, int head_dim) { return ((head_dim == 64 || head_dim == 128) && T >= 128 && T % 128 == 0) ? 1 : 0; }

extern "C" int oq_attn_fwd(const void* qkv, long ld, long width, int batch, int T, int heads, int kv_heads, int head_dim, int cq,
                           int gq, int ck, int cv, int gk, float alpha, void* O, long ldo, float* row_max, float* row_invl,
                           cudaStream_t stream) {
  if (check_common("oq_attn_fwd", qkv, ld, batch, T, heads, kv_heads, head_dim, cq, gq, ck, cv, gk, width)) return -1;
  const int D = head_dim;
  OQ_REQUIRE(O && row_max && row_invl && ldo % 8 == 0 && (uintptr_t)O % 16 == 0 && ldo >= (long)heads * D,
             "oq_attn_fwd: output / statistics pointers missing or misaligned");
  static bool attr = false;
  if (!attr) {
    if (int e = set_smem(attn_fwd_kernel<64>, fwd_smem<64>(), "oq_attn_fwd")) return e;
    if (int e = set_smem(attn_fwd_kernel<128>, fwd_smem<128>(), "oq_attn_fwd")) return e;
    attr = true;
  }
  CUtensorMap tm128;
  if (make_map(&tm128, qkv, width, (long)batch * T, ld, 128)) return -1;
  AttnParams p = base_params(batch, T, heads, kv_heads, D, cq, gq, ck, cv, gk, alpha);
  p.O = (bf16*)O; p.ldo = ldo; p.row_max = row_max; p.row_invl = row_invl;
  const int grid = batch * heads * (T / 128);
  if (D == 128) attn_fwd_kernel<128><<<grid, NUM_THREADS, fwd_smem<128>(), stream>>>(tm128, p);
  else attn_fwd_kernel<64><<<grid, NUM_THREADS, fwd_smem<64>(), stream>>>(tm128, p);
  return oq_check_launch("oq_attn_fwd");
}

extern "C" int oq_attn_bwd(const void* qkv, long ld, long width, int batch, int T, int heads, int kv_heads, int head_dim, int cq,
                           int gq, int ck, int cv, int gk, float alpha, const void* O, long ldo, const void* dO, long lddo,
                           const float* row_max, const float* row_invl, float* delta_ws, void* dqkv, long ldg,
                           cudaStream_t stream) {
  if (check_common("oq_attn_bwd", qkv, ld, batch, T, heads, kv_heads, head_dim, cq, gq, ck, cv, gk, width)) return -1;
  OQ_REQUIRE(O && dO && row_max && row_invl && delta_ws && dqkv, "oq_attn_bwd: null pointer");
  const int D = head_dim;
  OQ_REQUIRE(ldo % 8 == 0 && lddo % 8 == 0 && ldg % 8 == 0 && (uintptr_t)O % 16 == 0 && (uintptr_t)dO % 16 == 0 &&
                 (uintptr_t)dqkv % 16 == 0 && ldg >= width && lddo >= (long)heads * D && ldo >= (long)heads * D,
             "oq_attn_bwd: O / dO / dqkv must be 16-byte aligned with leading dimensions that are multiples of 8");
  static bool attr = false;
  if (!attr) {
    if (int e = set_smem(attn_dkv_kernel<64>, dkv_smem<64>(), "oq_attn_bwd")) return e;
    if (int e = set_smem(attn_dq_kernel<64>, dq_smem<64>(), "oq_attn_bwd")) return e;
    if (int e = set_smem(attn_dkv_kernel<128>, dkv_smem<128>(), "oq_attn_bwd")) return e;
    if (int e = set_smem(attn_dq_kernel<128>, dq_smem<128>(), "oq_attn_bwd")) return e;
    attr = true;
  }
  const long rows = (long)batch * T;
  CUtensorMap tm128, tm64, tmdo128, tmdo64;
  if (make_map(&tm128, qkv, width, rows, ld, 128) || make_map(&tm64, qkv, width, rows, ld, 64) ||
      make_map(&tmdo128, dO, (long)heads * D, rows, lddo, 128) || make_map(&tmdo64, dO, (long)heads * D, rows, lddo, 64))
    return -1;
  AttnParams p = base_params(batch, T, heads, kv_heads, D, cq, gq, ck, cv, gk, alpha);
  p.row_max = const_cast<float*>(row_max); p.row_invl = const_cast<float*>(row_invl); p.delta = delta_ws;
  p.dqkv = (bf16*)dqkv; p.ldg = ldg;
  const long nw = rows * heads;
  const unsigned gd = (unsigned)((nw + 7) / 8);
  const int gkv = batch * kv_heads * (T / 128), gq_ = batch * heads * (T / 128);
  if (D == 128) attn_delta_kernel<128><<<gd, 256, 0, stream>>>((const bf16*)O, ldo, (const bf16*)dO, lddo, delta_ws, batch, T, heads);
  else attn_delta_kernel<64><<<gd, 256, 0, stream>>>((const bf16*)O, ldo, (const bf16*)dO, lddo, delta_ws, batch, T, heads);
  if (int e = oq_check_launch("oq_attn_bwd (delta)")) return e;
  if (D == 128) attn_dkv_kernel<128><<<gkv, NUM_THREADS, dkv_smem<128>(), stream>>>(tm128, tm64, tmdo64, p);
  else attn_dkv_kernel<64><<<gkv, NUM_THREADS, dkv_smem<64>(), stream>>>(tm128, tm64, tmdo64, p);
  if (int e = oq_check_launch("oq_attn_bwd (dK, dV)")) return e;
  if (D == 128) attn_dq_kernel<128><<<gq_, NUM_THREADS, dq_smem<128>(), stream>>>(tm128, tm64, tmdo128, p);
  else attn_dq_kernel<64><<<gq_, NUM_THREADS, dq_smem<64>(), stream>>>(tm128, tm64, tmdo128, p);
  return oq_check_launch("oq_attn_bwd (dQ)");
}
