// DRAFT (never run on hardware -- see staging/README.md): fused causal attention forward for head_dim 128 on the already
// rotated + fake-quantised Q/K/V operands (QuantMatMul qkt_matmul + softmax + pv_matmul, reference
// models/int_llama_layer.py:141-163), DESIGN.md section 9 item 2.
//
// One CTA per (head, 128-query tile); key tiles j <= query tile (causal).  Two passes over the key tiles:
//   pass 1: S_j = Q K_j^T in TMEM -> row maximum m and l = sum exp(s - m) (online, registers only)
//   pass 2: S_j again -> P_j = bf16(exp(s - m) / l) with the FINAL statistics -> shared memory (UMMA A operand) and global
//           (the backward reads P) -> O += P_j V_j in TMEM
// so P is normalised before it is rounded, exactly like the unfused path (fp32 softmax, then the cast), and O never needs a
// correction.  Scores are scaled by alpha and rounded to fp16 before the softmax (the reference's fp16 autocast).
//
// warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM allocator   warps 4-7: softmax / epilogue (one query row per thread)
// TMEM columns: [0,128) S buffer 0, [128,256) S buffer 1, [256,384) O.   Shared memory: Q 32 KB | 2 x (K 32 KB + V 32 KB) | P 32 KB.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>

#include "../common.cuh"
#include "../mbar.cuh"

namespace {

constexpr int D = 128, BQ = 128, BKV = 128;
constexpr int BLK = 64 * 128 * 2;   // one 64-column K-major block of 128 rows: 16 KB (rows of 128 B, SWIZZLE_128B)
constexpr int TILE = 2 * BLK;       // a [128 x 128] bf16 tile: 32 KB
constexpr int KV_STAGES = 2;
constexpr int SMEM_BYTES = TILE /*Q*/ + KV_STAGES * 2 * TILE /*K,V*/ + TILE /*P*/ + 1024 /*align*/ + 256 /*barriers*/;
constexpr int NUM_THREADS = 256;

struct AttnParams {
  int T, heads;          // sequence length (multiple of 128), query heads
  int rep;               // query heads per K/V head (1 = MHA)
  int cq, ck, cv;        // column offsets of head 0's q / k / v inside the [T, ld] operand matrix
  float alpha;           // 1/sqrt(d)
  bf16* P;               // [heads, T, T] probabilities (for the backward)
  bf16* O; long ldo;     // [T, heads*D]
  float* row_max; float* row_sum;   // [heads, T] softmax statistics (for a recompute backward); may be null
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
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
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

// shared-memory matrix descriptor, SWIZZLE_128B (same encoding as gemm_sm100.cu::make_desc)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm128,   // box 64 cols x 128 rows (Q, K tiles: K-major operands)
                const __grid_constant__ CUtensorMap tm64,    // box 64 cols x  64 rows (V tiles: MN-major B operand)
                const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + TILE;                          // stage s: K at sKV + s*2*TILE, V at + TILE
  uint8_t* sP = sKV + KV_STAGES * 2 * TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + TILE);
  uint64_t* q_full = bars;                 // [1]
  uint64_t* kv_full = bars + 1;            // [2]
  uint64_t* kv_empty = bars + 3;           // [2]
  uint64_t* s_full = bars + 5;             // [2]  MMA -> softmax
  uint64_t* s_empty = bars + 7;            // [2]  softmax -> MMA (4 warps arrive)
  uint64_t* p_full = bars + 9;             // [1]  softmax -> MMA (4 warps arrive)
  uint64_t* p_empty = bars + 10;           // [1]  MMA (commit) -> softmax
  uint64_t* o_full = bars + 11;            // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_qt = p.T / BQ;
  const int qt = num_qt - 1 - (int)(blockIdx.x / p.heads);     // heavy (late) query tiles first
  const int head = (int)(blockIdx.x % p.heads);
  const int kvh = head / p.rep;
  const int nj = qt + 1;                                       // causal: key tiles 0..qt
  const int q_col = p.cq + head * D, k_col = p.ck + kvh * D, v_col = p.cv + kvh * D;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm128) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm64) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);
    }
    mbar_init(p_full, 4);
    mbar_init(p_empty, 1);
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
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, TILE);
      tma_load_2d(sQ, &tm128, q_full, q_col, qt * BQ);
      tma_load_2d(sQ + BLK, &tm128, q_full, q_col + 64, qt * BQ);
      int stage = 0;
      uint32_t phase = 0;
      for (int pass = 0; pass < 2; ++pass) {
        for (int j = 0; j < nj; ++j) {
          mbar_wait(&kv_empty[stage], phase ^ 1);
          uint8_t* sK = sKV + stage * 2 * TILE;
          uint8_t* sV = sK + TILE;
          mbar_expect_tx(&kv_full[stage], pass == 0 ? TILE : 2 * TILE);
          tma_load_2d(sK, &tm128, &kv_full[stage], k_col, j * BKV);
          tma_load_2d(sK + BLK, &tm128, &kv_full[stage], k_col + 64, j * BKV);
          if (pass == 1) {
            // V as MN-major B operand (n = head dim, k = key): per 64-key half, two 64-column blocks of [64 keys x 128 B]
#pragma unroll
            for (int kh = 0; kh < 2; ++kh)
#pragma unroll
              for (int nb = 0; nb < 2; ++nb)
                tma_load_2d(sV + kh * (2 * 64 * 128) + nb * (64 * 128), &tm64, &kv_full[stage], v_col + 64 * nb, j * BKV + 64 * kh);
          }
          if (++stage == KV_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // S = Q K^T: A, B K-major, M = 128, N = 128.   O += P V: A (P) K-major, B (V) MN-major, N = 128.
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BKV >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
      const uint32_t idesc_o = idesc_s | (1u << 16);
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
      mbar_wait(q_full, 0);
      tc_fence_after();
      int stage = 0, sb = 0;
      uint32_t kv_phase = 0, s_phase = 0, p_phase = 0;
      auto issue_s = [&](int st, int buf) {
        const uint32_t aK = smem_u32(sKV + st * 2 * TILE);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {        // 8 k-steps of 16 bf16: 4 per 64-column block
          const uint32_t off = (k >> 2) * BLK + (k & 3) * 32;
          tc_mma_bf16(tmem_base + buf * 128, make_desc(aQ + off, 16, 1024), make_desc(aK + off, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        }
      };
      // ---- pass 1: scores only
      for (int j = 0; j < nj; ++j) {
        mbar_wait(&kv_full[stage], kv_phase);
        mbar_wait(&s_empty[sb], s_phase ^ 1);
        tc_fence_after();
        issue_s(stage, sb);
        tc_commit(&kv_empty[stage]);
        tc_commit(&s_full[sb]);
        if (++stage == KV_STAGES) { stage = 0; kv_phase ^= 1; }
        if (++sb == 2) { sb = 0; s_phase ^= 1; }
      }
      // ---- pass 2: scores, then O += P V once the softmax warps have written P
      for (int j = 0; j < nj; ++j) {
        mbar_wait(&kv_full[stage], kv_phase);
        mbar_wait(&s_empty[sb], s_phase ^ 1);
        tc_fence_after();
        issue_s(stage, sb);
        tc_commit(&s_full[sb]);
        mbar_wait(p_full, p_phase);
        tc_fence_after();
        const uint32_t aV = smem_u32(sKV + stage * 2 * TILE + TILE);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {      // 8 k-steps over the 128 keys: 4 per 64-key half
          const uint32_t a_off = (k >> 2) * BLK + (k & 3) * 32;                 // P: K-major blocks of 64 keys
          const uint32_t b_off = (k >> 2) * (2 * 64 * 128) + (k & 3) * (16 * 128);   // V: 16 key rows of 128 B per k-step
          tc_mma_bf16(tmem_base + 256, make_desc(aP + a_off, 16, 1024), make_desc(aV + b_off, 64 * 128, 1024), idesc_o,
                      (j > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(p_empty);
        tc_commit(&kv_empty[stage]);
        p_phase ^= 1;
        if (++stage == KV_STAGES) { stage = 0; kv_phase ^= 1; }
        if (++sb == 2) { sb = 0; s_phase ^= 1; }
      }
      tc_commit(o_full);
    }
  } else if (warp >= 4) {
    // ===================== softmax / epilogue: thread = one query row =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;                 // row inside the tile
    const int qrow = qt * BQ + r;                // query position
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float m = -INFINITY, l = 0.f;
    int sb = 0;
    uint32_t s_phase = 0, pe_phase = 0;
    auto score = [&](uint32_t bits) {            // alpha * acc, rounded to fp16 like the reference's autocast
      return __half2float(__float2half_rn(__uint_as_float(bits) * p.alpha));
    };
    // ---- pass 1
    for (int j = 0; j < nj; ++j) {
      mbar_wait(&s_full[sb], s_phase);
      tc_fence_after();
      const int lim = (j == qt) ? r + 1 : BKV;   // causal: keys c <= r on the diagonal tile
      for (int c0 = 0; c0 < BKV; c0 += 32) {
        uint32_t v[32];
        tc_ld32(lane_base + sb * 128 + c0, v);
        tc_wait_ld();
        float s[32], mx = m;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          s[i] = (c0 + i < lim) ? score(v[i]) : -INFINITY;
          mx = fmaxf(mx, s[i]);
        }
        float add = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) add += __expf(s[i] - mx);
        l = l * __expf(m - mx) + add;            // m = -inf, l = 0 on the first chunk: exp(-inf) = 0
        m = mx;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);
      if (++sb == 2) { sb = 0; s_phase ^= 1; }
    }
    const float inv = 1.0f / l;
    if (p.row_max) {
      p.row_max[(long)head * p.T + qrow] = m;
      p.row_sum[(long)head * p.T + qrow] = l;
    }
    // ---- pass 2
    bf16* prow = p.P + ((long)head * p.T + qrow) * p.T;
    for (int j = 0; j < nj; ++j) {
      mbar_wait(&s_full[sb], s_phase);
      tc_fence_after();
      if (j > 0) {                               // the previous P V MMA must have finished reading sP
        mbar_wait(p_empty, pe_phase);
        pe_phase ^= 1;
      }
      const int lim = (j == qt) ? r + 1 : BKV;
      for (int c0 = 0; c0 < BKV; c0 += 32) {
        uint32_t v[32];
        tc_ld32(lane_base + sb * 128 + c0, v);
        tc_wait_ld();
        float pr[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) pr[i] = (c0 + i < lim) ? __expf(score(v[i]) - m) * inv : 0.f;
        // shared memory (UMMA A operand, K-major, SWIZZLE_128B): 64-key block c0/64, row r, 16-byte chunk ^ (r & 7)
        uint8_t* blk = sP + (c0 >> 6) * BLK + r * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint4 u;
          u.x = pack_bf2(pr[ch * 8 + 0], pr[ch * 8 + 1]); u.y = pack_bf2(pr[ch * 8 + 2], pr[ch * 8 + 3]);
          u.z = pack_bf2(pr[ch * 8 + 4], pr[ch * 8 + 5]); u.w = pack_bf2(pr[ch * 8 + 6], pr[ch * 8 + 7]);
          const int chunk = ((c0 & 63) >> 3) + ch;                         // 16-byte chunk inside the 128-byte row
          *reinterpret_cast<uint4*>(blk + ((chunk ^ (r & 7)) << 4)) = u;
          *reinterpret_cast<uint4*>(prow + j * BKV + c0 + ch * 8) = u;     // global copy for the backward
        }
      }
      fence_proxy_async();                       // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(p_full);
        mbar_arrive(&s_empty[sb]);
      }
      if (++sb == 2) { sb = 0; s_phase ^= 1; }
    }
    // ---- epilogue: O
    mbar_wait(o_full, 0);
    tc_fence_after();
    bf16* orow = p.O + (long)qrow * p.ldo + head * D;
    for (int c0 = 0; c0 < D; c0 += 32) {
      uint32_t v[32];
      tc_ld32(lane_base + 256 + c0, v);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        float t[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = __uint_as_float(v[i + e]);
        store_bf8(orow + c0 + i, t);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace

// Host side (to write when the kernel moves into csrc/): two 2-D bf16 tensor maps over the [T, ld] operand matrix
// (boxes 64 x 128 and 64 x 64, SWIZZLE_128B), cudaFuncSetAttribute(MaxDynamicSharedMemorySize = SMEM_BYTES),
// grid = heads * (T / 128), block = 256.
