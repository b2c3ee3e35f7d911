// Register-resident fast paths of the weight fake-quant forward / backward (included by weight_quant.cu).
//
// Per-channel (group_size 0): one 256-thread block per output row, the whole row held in registers as NV
// float4 per thread (K <= 1024*NV), all loads of a row issued back to back and the next row prefetched, one
// block barrier per row.  Grouped: LPG lanes x VPL float4 per lane hold one group (group = 4*LPG*VPL), a warp
// holds 32/LPG groups at once.  Same arithmetic (oq_math.h) as the generic kernels; the division inside
// round(x/scale) goes through oq_rint_div (bit-identical, divides only near rounding boundaries).
#pragma once

// ------------------------------------------------------------------ block reduction with one barrier
// Each thread contributes NS sums and NM maxima; results broadcast.  `scratch` holds 2*(NS+NM)*8 floats and
// is double-buffered on `par` so that no trailing barrier is needed between consecutive rows.
template <int NS, int NM>
__device__ __forceinline__ void block_reduce_1sync(float (&s)[NS > 0 ? NS : 1], float (&m)[NM > 0 ? NM : 1], float* scratch,
                                                   int par) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NS; ++i) s[i] = warp_reduce(s[i], OpSum());
#pragma unroll
  for (int i = 0; i < NM; ++i) m[i] = warp_reduce(m[i], OpMax());
  float* buf = scratch + par * (NS + NM) * 8;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NS; ++i) buf[i * 8 + warp] = s[i];
#pragma unroll
    for (int i = 0; i < NM; ++i) buf[(NS + i) * 8 + warp] = m[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += buf[i * 8 + w];
    s[i] = t;
  }
#pragma unroll
  for (int i = 0; i < NM; ++i) {
    float t = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) t = fmaxf(t, buf[(NS + i) * 8 + w]);
    m[i] = t;
  }
}

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float& f4(float4& v, int j) { return reinterpret_cast<float*>(&v)[j]; }

// ================================================================== per-channel forward
// warp-wide fp32 max in ONE instruction (sm_100a: CREDUX.MAX.F32); NaN operands are ignored like fmaxf
__device__ __forceinline__ float warp_max_f32(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

// Instruction-count notes (ncu, K=4096: the first version executed 50 thread-instructions per weight, 10 % of them
// branches on run-time-uniform LET options inside the element loops): the row-division fold is a template parameter,
// absent column scale / shift vectors are read as 1 / 0 from shared memory (w*1 and +w*0 are exact), the sigmoids of
// all rows a block visits are evaluated once up front (one row per thread), the block-wide min / max is two CREDUX
// + one barrier, and the bit-exact rounding guard of oq_rint_div is evaluated branch-free with one fallback branch
// per float4 (the exact IEEE-division path re-does those four elements; probability ~1e-5).
// ROWMODE: 0 the row fold multiplies (or is absent), 1 it divides, 2 per-row segment table (q | k | v in one launch)
template <int NV, bool LET, int ROWMODE>
__global__ void __launch_bounds__(256) wq_pc_fwd(WqArgs a) {
  extern __shared__ float4 dsm4[];
  __shared__ float red[2][4][8];
  __shared__ float sig[2][256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, K4 = a.K >> 2;
  float4* colv = dsm4;            // [K4] (LET)
  float4* shv = dsm4 + K4;        // [K4] (LET)
  const bool has_shift = LET && a.shift != nullptr;
  if (LET) {
    for (int i = tid; i < K4; i += 256) {
      colv[i] = a.col ? ld4(a.col + 4 * i) : make_float4(1.f, 1.f, 1.f, 1.f);
      shv[i] = a.shift ? ld4(a.shift + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const int lwc = a.up != nullptr;
  float4 wn[NV];
  int n = blockIdx.x;
  if (n < a.N) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = tid + i * 256;
      wn[i] = idx < K4 ? ld4(a.W + (long)n * a.K + 4 * idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  int par = 0, k = 0;
  int pend_n = -1;          // s8 side output: the row whose code sum is still in red[par ^ 1][3] (thread 0 only)
  float pend_scale = 0.f;
  for (; n < a.N; n += gridDim.x, par ^= 1, ++k) {
    if ((k & 255) == 0) {   // sigmoids of the next 256 rows of this block, one row per thread
      __syncthreads();
      const long r = (long)n + (long)tid * gridDim.x;
      const bool ok = lwc && r < a.N;
      sig[0][tid] = ok ? oq_sigmoidf(a.up[r]) : 1.0f;
      sig[1][tid] = ok ? oq_sigmoidf(a.low[r]) : 1.0f;
      __syncthreads();
    }
    float4 w[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) w[i] = wn[i];
    const int nn = n + gridDim.x;
    if (nn < a.N) {  // prefetch the next row while this one is processed
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int idx = tid + i * 256;
        if (idx < K4) wn[i] = ld4(a.W + (long)nn * a.K + 4 * idx);
      }
    }
    float rs = 1.0f;
    bool rdiv = ROWMODE == 1;
    if (LET) {
      if (ROWMODE == 2) {
        int op;
        wq_row_fold(a, n, rs, op);
        rdiv = op == 2;
      } else {
        rs = (a.row && a.row_op != 0) ? a.row[n] : 1.0f;
      }
    }
    float sm = 0.f, mlo = -INFINITY, mhi = -INFINITY;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = tid + i * 256;
      if (idx < K4) {
        if (LET) {
          const float4 c4 = colv[idx], s4 = shv[idx];
          sm += w[i].x * s4.x + w[i].y * s4.y + w[i].z * s4.z + w[i].w * s4.w;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float f1 = OQ_MUL(f4(w[i], j), f4(const_cast<float4&>(c4), j));
            f4(w[i], j) = rdiv ? OQ_DIV(f1, rs) : OQ_MUL(f1, rs);  // (W * col) {*,/} row; a factor 1 is exact
          }
        }
        mlo = fmaxf(mlo, fmaxf(fmaxf(-w[i].x, -w[i].y), fmaxf(-w[i].z, -w[i].w)));
        mhi = fmaxf(mhi, fmaxf(fmaxf(w[i].x, w[i].y), fmaxf(w[i].z, w[i].w)));
      }
    }
    mlo = warp_max_f32(mlo);
    mhi = warp_max_f32(mhi);
    if (has_shift) sm = warp_reduce(sm, OpSum());
    if (lane == 0) {
      red[par][0][warp] = mlo;
      red[par][1][warp] = mhi;
      red[par][2][warp] = sm;
    }
    __syncthreads();
    const float xmin = -warp_max_f32(lane < 8 ? red[par][0][lane] : -INFINITY);
    const float xmax = warp_max_f32(lane < 8 ? red[par][1][lane] : -INFINITY);
    const OqQParams q = oq_qparams(xmin, xmax, sig[0][k & 255], sig[1][k & 255], lwc, a.n_bits, a.symmetric, a.recip_div);
    const float z0 = oq_s8_z0(q.zp, q.qmax);
    if (tid == 0) {
      if (a.codes8) {
        if (pend_n >= 0) {     // the previous row's warp sums were published before this row's barrier
          float t = 0.f;
#pragma unroll
          for (int x = 0; x < 8; ++x) t += red[par ^ 1][3][x];
          a.s8_t[pend_n] = pend_scale * t;
        }
        pend_n = n;
        pend_scale = q.scale;
        // a degenerate (constant) row dequantises to NaN in the fake-quant path: poison the epilogue term instead
        a.s8_beta[n] = isfinite(s8_scale_or_nan(q)) && q.scale != 0.0f ? (z0 - q.zp) * q.scale : __int_as_float(0x7fc00000);
      }
      if (a.scale_out) a.scale_out[n] = q.scale;
      if (a.zp_out) a.zp_out[n] = q.zp;
      if (a.xmin_out) a.xmin_out[n] = xmin;
      if (a.xmax_out) a.xmax_out[n] = xmax;
      if (has_shift && a.mv) {
        float t = 0.f;
#pragma unroll
        for (int x = 0; x < 8; ++x) t += red[par][2][x];
        a.mv[n] = t;
      }
    }
    const float inv = OQ_DIV(1.0f, q.scale);
    const bool inv_bad = !oq_inv_ok(inv);
    int isum = 0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = tid + i * 256;
      if (idx < K4) {
        float4 c;
        bool redo = inv_bad;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float qa = OQ_MUL(f4(w[i], j), inv);
          const float r = rintf(qa);
          redo |= !(fabsf(fabsf(qa - r) - 0.5f) > fabsf(qa) * 2.4e-7f + 1e-30f);   // oq_rint_div's guard
          f4(c, j) = fminf(fmaxf(OQ_ADD(r, q.zp), 0.0f), q.qmax);    // finite here: scale is normal, weights finite
        }
        if (redo) {   // exact IEEE division (and NaN-preserving clamp) for this float4
#pragma unroll
          for (int j = 0; j < 4; ++j) f4(c, j) = oq_code(f4(w[i], j), q);
        }
        float y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) y[j] = oq_dequant(f4(c, j), q);
        const long o = (long)n * a.K + 4 * idx;
        if (a.q_bf) store_bf4(a.q_bf + (long)n * a.ld_bf + 4 * idx, y);
        if (a.q_f32) *reinterpret_cast<float4*>(a.q_f32 + o) = make_float4(y[0], y[1], y[2], y[3]);
        if (a.codes)
          *reinterpret_cast<uchar4*>(a.codes + o) = make_uchar4((unsigned char)c.x, (unsigned char)c.y, (unsigned char)c.z, (unsigned char)c.w);
        if (a.codes8) {
          const float cc[4] = {c.x, c.y, c.z, c.w};
          s8_store4(a.codes8 + o, cc, z0, isum);
        }
      }
    }
    if (a.codes8) {
      const int wsum = __reduce_add_sync(0xffffffffu, isum);      // one instruction (REDUX) instead of five shuffles
      if (lane == 0) red[par][3][warp] = (float)wsum;
    }
  }
  if (a.codes8) {
    __syncthreads();
    if (tid == 0 && pend_n >= 0) {
      float t = 0.f;
#pragma unroll
      for (int x = 0; x < 8; ++x) t += red[par ^ 1][3][x];
      a.s8_t[pend_n] = pend_scale * t;
    }
  }
}

// ------------------------------------------------------------------ per-channel forward, calibration-loop variant
// The optimiser step calls the forward with ONE output format (bf16 W_q, plus the s8 side output when the forward GEMM
// is the exact-integer one).  ncu on the general kernel above (profiles/r2_wq_ncu_summary.csv): 43.6 thread-instructions
// per weight at K = 4096, of which ~17 are integer / control (null checks of the optional outputs and bounds tests per
// float4, 64-bit index arithmetic) and ~13 the quantise step.  This variant removes them:
//  * outputs fixed by template (S8), bounds test only on a thread's LAST float4, row pointers hoisted;
//  * clamp BEFORE the rounding: clamp(rint(qa) + zp, 0, qmax) == rint(clamp(qa, -zp, qmax - zp)) + zp because zp is an
//    integer and rint is monotone and fixes integers.  The clamped quotient is bounded, so rint is the magic-number add
//    t = qc + M (M = 1.5 * 2^23 + (zp - z0): the low byte of t's bit pattern IS the s8 code), r = t - M is the
//    dequantised integer (code - zp, exact), W_q = r * scale -- the same fp32 product as oq_dequant;
//  * the bit-exactness guard of oq_rint_div in the clamped domain: qa = f * (1/scale) is within 1.8e-7 |q| of the
//    IEEE quotient, so rint can only differ when |qc - r| is that close to 0.5; |q| <= max(|zp|, |qmax - zp|) + 1
//    inside the clamp range gives a ROW-constant threshold, and the four |qc - r| of a float4 are folded with one
//    3-input max (2.5 instructions per weight instead of 5).  A failing float4 (or a row whose 1/scale is not a normal
//    number) is redone with the IEEE division (oq_code / oq_dequant);
//  * the row extrema use the NaN-propagating 3-input min / max, so a NaN / Inf weight makes the row's scale non-finite
//    and sends the whole row through the exact path (the reference's amin / amax propagate NaN the same way);
//  * the q | k | v row table is resolved once per THREADS rows (with the sigmoids), not per row by every thread.
// ROWMODE: 0 multiply by row[n], 1 divide, 2 per-row segment table, 3 no row fold.
__device__ __forceinline__ float fmax3_nan(float a, float b, float c) {
  float d;
  asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float fmin3_nan(float a, float b, float c) {
  float d;
  asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float warp_max_nan(float v) {
  float r;
  asm volatile("redux.sync.max.NaN.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float warp_min_nan(float v) {
  float r;
  asm volatile("redux.sync.min.NaN.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

template <int NV, int THREADS, bool LET, int ROWMODE, bool S8, bool EXACT>
__global__ void __launch_bounds__(THREADS) wq_pc_fwd_lean(WqArgs a) {
  extern __shared__ float4 dsm4[];
  constexpr int NW = THREADS / 32;
  static_assert(NW == 4 || NW == 8, "128 or 256 threads");
  __shared__ __align__(16) float red[2][4][NW];
  __shared__ float sig[2][THREADS];
  __shared__ float rsv[THREADS];
  __shared__ int rdv[THREADS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, K4 = a.K >> 2;
  float4* colv = dsm4;            // [K4] (LET)
  float4* shv = dsm4 + K4;        // [K4] (LET)
  const bool has_shift = LET && a.shift != nullptr;
  if (LET) {
    for (int i = tid; i < K4; i += THREADS) {
      colv[i] = a.col ? ld4(a.col + 4 * i) : make_float4(1.f, 1.f, 1.f, 1.f);
      shv[i] = a.shift ? ld4(a.shift + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const int lwc = a.up != nullptr;
  // EXACT: (NV - 1) * THREADS < K4 <= NV * THREADS, only a thread's last float4 can fall off the row (everything else is
  // resolved at compile time); otherwise a thread owns its first `nfull` float4s
  const bool last_ok = tid + (NV - 1) * THREADS < K4;
  const int nfull = (K4 - tid + THREADS - 1) / THREADS;
#define OQ_LEAN_OK(i) (EXACT ? ((i) < NV - 1 || last_ok) : ((i) < nfull))
  const float* wbase = a.W + 4 * tid;
  float4 wn[NV];
  int n = blockIdx.x;
  if (n < a.N) {
    const float* wr = wbase + (long)n * a.K;
#pragma unroll
    for (int i = 0; i < NV; ++i) wn[i] = OQ_LEAN_OK(i) ? ld4(wr + 4 * i * THREADS) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int par = 0, k = 0;
  int pend_n = -1;          // s8 side output: the row whose code sum is still in red[par ^ 1][3] (thread 0 only)
  float pend_scale = 0.f;
  for (; n < a.N; n += gridDim.x, par ^= 1, ++k) {
    if (k == THREADS) k = 0;
    if (k == 0) {   // per-row scalars of the next THREADS rows of this block, one row per thread
      __syncthreads();
      const long r = (long)n + (long)tid * gridDim.x;
      const bool ok = r < a.N;
      sig[0][tid] = (ok && lwc) ? oq_sigmoidf(a.up[r]) : 1.0f;
      sig[1][tid] = (ok && lwc) ? oq_sigmoidf(a.low[r]) : 1.0f;
      if (LET && ROWMODE != 3) {
        float rs = 1.0f;
        int op = 0;
        if (ok) wq_row_fold(a, r, rs, op);
        rsv[tid] = rs;
        rdv[tid] = op == 2;
      }
      __syncthreads();
    }
    float4 w[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) w[i] = wn[i];
    const int nn = n + gridDim.x;
    if (nn < a.N) {  // prefetch the next row while this one is processed
      const float* wr = wbase + (long)nn * a.K;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if OQ_LEAN_OK(i) wn[i] = ld4(wr + 4 * i * THREADS);
    }
    // ---- fold + extrema
    float sm = 0.f, mlo = INFINITY, mhi = -INFINITY;
    if (LET) {
      const float rs = ROWMODE != 3 ? rsv[k] : 1.0f;
      const bool rdiv = ROWMODE == 1 || (ROWMODE == 2 && rdv[k]);
      if (rdiv) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          if OQ_LEAN_OK(i) {
            const float4 c4 = colv[tid + i * THREADS], s4 = shv[tid + i * THREADS];
            sm += w[i].x * s4.x + w[i].y * s4.y + w[i].z * s4.z + w[i].w * s4.w;
#pragma unroll
            for (int j = 0; j < 4; ++j) f4(w[i], j) = OQ_DIV(OQ_MUL(f4(w[i], j), f4(const_cast<float4&>(c4), j)), rs);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          if OQ_LEAN_OK(i) {
            const float4 c4 = colv[tid + i * THREADS], s4 = shv[tid + i * THREADS];
            sm += w[i].x * s4.x + w[i].y * s4.y + w[i].z * s4.z + w[i].w * s4.w;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float f1 = OQ_MUL(f4(w[i], j), f4(const_cast<float4&>(c4), j));
              f4(w[i], j) = ROWMODE == 3 ? f1 : OQ_MUL(f1, rs);     // (W * col) * row; a factor 1 is exact
            }
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if OQ_LEAN_OK(i) {
        mlo = fmin3_nan(fmin3_nan(mlo, w[i].x, w[i].y), w[i].z, w[i].w);
        mhi = fmax3_nan(fmax3_nan(mhi, w[i].x, w[i].y), w[i].z, w[i].w);
      }
    }
    mlo = warp_min_nan(mlo);
    mhi = warp_max_nan(mhi);
    if (has_shift) sm = warp_reduce(sm, OpSum());
    if (lane == 0) {
      red[par][0][warp] = mlo;
      red[par][1][warp] = mhi;
      red[par][2][warp] = sm;
    }
    __syncthreads();
    float xmin, xmax;
    {
      const float4 a0 = *reinterpret_cast<const float4*>(&red[par][0][0]), b0 = *reinterpret_cast<const float4*>(&red[par][1][0]);
      xmin = fmin3_nan(fmin3_nan(a0.x, a0.y, a0.z), a0.w, a0.w);
      xmax = fmax3_nan(fmax3_nan(b0.x, b0.y, b0.z), b0.w, b0.w);
      if (NW == 8) {
        const float4 a1 = *reinterpret_cast<const float4*>(&red[par][0][4]), b1 = *reinterpret_cast<const float4*>(&red[par][1][4]);
        xmin = fmin3_nan(fmin3_nan(xmin, a1.x, a1.y), a1.z, a1.w);
        xmax = fmax3_nan(fmax3_nan(xmax, b1.x, b1.y), b1.z, b1.w);
      }
    }
    const OqQParams q = oq_qparams(xmin, xmax, sig[0][k], sig[1][k], lwc, a.n_bits, a.symmetric, a.recip_div);
    const float z0 = S8 ? oq_s8_z0(q.zp, q.qmax) : q.zp;
    if (tid == 0) {
      if (S8) {
        if (pend_n >= 0) {     // the previous row's warp sums were published before this row's barrier
          float t = 0.f;
#pragma unroll
          for (int x = 0; x < NW; ++x) t += red[par ^ 1][3][x];
          a.s8_t[pend_n] = pend_scale * t;
        }
        pend_n = n;
        pend_scale = q.scale;
        // a degenerate (constant) row dequantises to NaN in the fake-quant path: poison the epilogue term instead
        a.s8_beta[n] = isfinite(s8_scale_or_nan(q)) && q.scale != 0.0f ? (z0 - q.zp) * q.scale : __int_as_float(0x7fc00000);
      }
      if (a.scale_out) a.scale_out[n] = q.scale;
      if (a.zp_out) a.zp_out[n] = q.zp;
      if (a.xmin_out) a.xmin_out[n] = xmin;
      if (a.xmax_out) a.xmax_out[n] = xmax;
      if (has_shift && a.mv) {
        float t = 0.f;
#pragma unroll
        for (int x = 0; x < NW; ++x) t += red[par][2][x];
        a.mv[n] = t;
      }
    }
    // ---- quantise
    const float inv = OQ_DIV(1.0f, q.scale);
    const float lo = -q.zp, hi = q.qmax - q.zp;                     // integers: exact
    const float magic = 12582912.0f + (q.zp - z0);                 // S8: low byte of (qc + magic) = code - z0
    const bool fast_ok = oq_inv_ok(inv) && fabsf(q.zp) <= 1.0e4f;  // (a NaN zero point fails the comparison)
    const float thr = fast_ok ? 0.5f - ((fmaxf(fabsf(lo), fabsf(hi)) + 1.0f) * 2.4e-7f + 1e-30f) : -1.0f;
    bf16* qrow = a.q_bf + (long)n * a.ld_bf + 4 * tid;
    int8_t* crow = S8 ? a.codes8 + (long)n * a.K + 4 * tid : nullptr;
    int isum = 0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if OQ_LEAN_OK(i) {
        float t[4], y[4], d[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float qa = OQ_MUL(f4(w[i], j), inv);
          const float qc = fminf(fmaxf(qa, lo), hi);
          t[j] = OQ_ADD(qc, magic);
          const float r = OQ_SUB(t[j], magic);
          d[j] = OQ_SUB(qc, r);
          y[j] = OQ_MUL(r, q.scale);
        }
        const float m = fmaxf(fmaxf(fabsf(d[0]), fabsf(d[1])), fmaxf(fabsf(d[2]), fabsf(d[3])));
        if (!(m < thr)) {   // exact IEEE division (and NaN-preserving clamp) for this float4
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float c = oq_code(f4(w[i], j), q);
            y[j] = oq_dequant(c, q);
            t[j] = (c - z0) + 12582912.0f;
          }
        }
        store_bf4(qrow + 4 * i * THREADS, y);
        if (S8) {
          const uint32_t w8 = __byte_perm(__byte_perm(__float_as_uint(t[0]), __float_as_uint(t[1]), 0x0040),
                                          __byte_perm(__float_as_uint(t[2]), __float_as_uint(t[3]), 0x0040), 0x5410);
          isum = __dp4a((int)w8, 0x01010101, isum);
          *reinterpret_cast<uint32_t*>(crow + 4 * i * THREADS) = w8;
        }
      }
    }
    if (S8) {
      const int wsum = __reduce_add_sync(0xffffffffu, isum);      // one instruction (REDUX) instead of five shuffles
      if (lane == 0) red[par][3][warp] = (float)wsum;
    }
  }
  if (S8) {
    __syncthreads();
    if (tid == 0 && pend_n >= 0) {
      float t = 0.f;
#pragma unroll
      for (int x = 0; x < NW; ++x) t += red[par ^ 1][3][x];
      a.s8_t[pend_n] = pend_scale * t;
    }
  }
#undef OQ_LEAN_OK
}

// ================================================================== per-channel backward
// dynamic smem (LET): col[K] | colacc[K] | shacc[K]
template <int NV, bool LET>
__global__ void __launch_bounds__(256) wq_pc_bwd(WqArgs a) {
  extern __shared__ float4 dsm4[];
  __shared__ float scratch[2 * 6 * 8];
  const int tid = threadIdx.x, K4 = a.K >> 2;
  float4* colv = dsm4;
  float4* colacc = dsm4 + K4;
  float4* shacc = dsm4 + 2 * K4;
  const bool has_col = LET && a.col != nullptr;
  const bool want_col = LET && a.dcol != nullptr;
  const bool want_sh = LET && a.dshift != nullptr;
  if (LET) {
    for (int i = tid; i < K4; i += 256) {
      if (has_col) colv[i] = ld4(a.col + 4 * i);
      colacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      shacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
  }
  const int lwc = a.up != nullptr;
  int par = 0;
  auto load_row = [&](int row, float4 (&ww)[NV], float4 (&gg)[NV]) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = tid + i * 256;
      if (idx < K4) {
        ww[i] = ld4(a.W + (long)row * a.K + 4 * idx);
        if (a.g_is_f32) {
          gg[i] = ld4(reinterpret_cast<const float*>(a.g) + (long)row * a.ld_g + 4 * idx);
        } else {
          float t[4];
          load_bf4(reinterpret_cast<const bf16*>(a.g) + (long)row * a.ld_g + 4 * idx, t);
          gg[i] = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
    }
  };
  constexpr bool PREFETCH = false;   // measured: the extra registers (3 -> 2 blocks/SM) cost more than the overlap wins
  float4 wn[PREFETCH ? NV : 1], gn[PREFETCH ? NV : 1];
  if constexpr (PREFETCH) {
    if ((int)blockIdx.x < a.N) load_row(blockIdx.x, wn, gn);
  }
  for (int n = blockIdx.x; n < a.N; n += gridDim.x, par ^= 1) {
    float4 w[NV], g[NV];
    if constexpr (PREFETCH) {
#pragma unroll
      for (int i = 0; i < NV; ++i) { w[i] = wn[i]; g[i] = gn[i]; }
      if (n + (int)gridDim.x < a.N) load_row(n + gridDim.x, wn, gn);
    } else {
      load_row(n, w, g);
    }
    const float rs = (LET && a.row) ? a.row[n] : 1.0f;
    const int rop = (LET && a.row) ? a.row_op : 0;
    const float rfac = rop == 1 ? rs : (rop == 2 ? 1.0f / rs : 1.0f);
    const float en = (LET && a.e_row) ? a.e_row[n] : 0.0f;
    const float xmin = a.xmin_in[n], xmax = a.xmax_in[n];
    OqQParams q;
    if (a.scale_in) {  // saved by the forward: skips two sigmoids and the scale / zero-point divisions per row
      q.scale = a.scale_in[n];
      q.zp = a.zp_in[n];
      q.qmax = (float)((1 << a.n_bits) - 1);
    } else {
      q = oq_qparams(xmin, xmax, lwc ? oq_sigmoidf(a.up[n]) : 1.0f, lwc ? oq_sigmoidf(a.low[n]) : 1.0f, lwc, a.n_bits,
                     a.symmetric, a.recip_div);
    }
    const float inv = OQ_DIV(1.0f, q.scale);
    const int inv_ok = oq_inv_ok(inv);
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // R, #max, #min, rowdot, sum w1@max, sum w1@min
    float dummy[1] = {-INFINITY};
    bool tie = false;
    // The backward is tolerance-checked (not bit-exact), so the fold uses a reciprocal multiply and the arg-max /
    // arg-min elements are recognised within 2 ulp of the saved extrema (which the forward computed exactly).
    const float rmul = rop == 1 ? rs : (rop == 2 ? 1.0f / rs : 1.0f);
    const float tie_tol_max = 2.4e-7f * fabsf(xmax), tie_tol_min = 2.4e-7f * fabsf(xmin);
    const bool tie_all = !(lwc && xmax > 0.0f && xmin < 0.0f) || a.symmetric;
    const float qmid = 0.5f * q.qmax, qhalf_m1 = 0.5f * q.qmax - 1.0f;   // |t - Q/2| >= Q/2 - 1  <=>  t <= 1 or t >= Q-1
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = tid + i * 256;
      if (idx < K4) {
        const float4 c4 = has_col ? colv[idx] : make_float4(1.f, 1.f, 1.f, 1.f);
        float4 ca = make_float4(0.f, 0.f, 0.f, 0.f), sa = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float wv = f4(w[i], j);
          const float w1 = has_col ? wv * f4(const_cast<float4&>(c4), j) : wv;
          const float f = w1 * rmul;          // (W*col) {*,/} row with the division as a reciprocal multiply
          const float qa = f * inv;
          const float t = rintf(qa) + q.zp;
          const float c = fminf(fmaxf(t, 0.0f), q.qmax);      // (NaN rows are not representable here; the forward owns them)
          const bool inr = (c == t);
          const float gk = f4(g[i], j);
          acc[0] += gk * ((c - q.zp) - (inr ? qa : 0.0f));
          if (tie_all || fabsf(t - qmid) >= qhalf_m1) {   // with an LWC clip (sigmoid < 1) the extrema land on the outermost codes (+-1 from the double rounding of scale and zero point)
            const bool ismax = fabsf(f - xmax) <= tie_tol_max, ismin = fabsf(f - xmin) <= tie_tol_min;
            if (ismax) { acc[1] += 1.0f; acc[4] += w1; }
            if (ismin) { acc[2] += 1.0f; acc[5] += w1; }
            tie |= ismax | ismin;
          }
          const float d2 = inr ? gk : 0.0f;
          acc[3] += d2 * w1;
          f4(ca, j) = d2 * rfac * wv;
          f4(sa, j) = wv * en;
          f4(g[i], j) = d2;  // keep the straight-through part for the optional full-gradient output
        }
        if (want_col) {
          float4 t = colacc[idx];
          t.x += ca.x; t.y += ca.y; t.z += ca.z; t.w += ca.w;
          colacc[idx] = t;
        }
        if (want_sh) {
          float4 t = shacc[idx];
          t.x += sa.x; t.y += sa.y; t.z += sa.z; t.w += sa.w;
          shacc[idx] = t;
        }
      }
    }
    block_reduce_1sync<6, 0>(acc, dummy, scratch, par);
    float cmax = 0.f, cmin = 0.f;
    if (tid == 0 || tie || a.dW) {   // only the threads that need the clip-path constants evaluate the sigmoids
      const float su = lwc ? oq_sigmoidf(a.up[n]) : 1.0f;
      const float sl = lwc ? oq_sigmoidf(a.low[n]) : 1.0f;
      float gmaxp, gminp;
      oq_scale_grad(acc[0], xmin, xmax, su, sl, lwc, a.n_bits, a.symmetric, a.recip_div, &gmaxp, &gminp);
      cmax = acc[1] > 0.f ? gmaxp * su / acc[1] : 0.0f;
      cmin = acc[2] > 0.f ? gminp * sl / acc[2] : 0.0f;
      if (tid == 0) {
        if (lwc && a.d_up) a.d_up[n] = gmaxp * xmax * su * (1.0f - su);
        if (lwc && a.d_low) a.d_low[n] = gminp * xmin * sl * (1.0f - sl);
        if (a.drow) a.drow[n] = acc[3] + cmax * acc[4] + cmin * acc[5];
      }
    }
    if (tie || a.dW) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int idx = tid + i * 256;
        if (idx < K4) {
          const float4 c4 = has_col ? colv[idx] : make_float4(1.f, 1.f, 1.f, 1.f);
          float4 corr4 = make_float4(0.f, 0.f, 0.f, 0.f);
          bool any = false;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float wv = f4(w[i], j);
            const float w1 = has_col ? wv * f4(const_cast<float4&>(c4), j) : wv;
            const float f = w1 * rmul;
            float corr = 0.0f;
            if (fabsf(f - xmax) <= tie_tol_max) corr += cmax;
            if (fabsf(f - xmin) <= tie_tol_min) corr += cmin;
            f4(corr4, j) = corr;
            any |= corr != 0.0f;
          }
          if (any && want_col) {
            float4 t = colacc[idx];
            t.x += corr4.x * rfac * w[i].x; t.y += corr4.y * rfac * w[i].y;
            t.z += corr4.z * rfac * w[i].z; t.w += corr4.w * rfac * w[i].w;
            colacc[idx] = t;
          }
          if (a.dW)
            *reinterpret_cast<float4*>(a.dW + (long)n * a.K + 4 * idx) =
                make_float4(g[i].x + corr4.x, g[i].y + corr4.y, g[i].z + corr4.z, g[i].w + corr4.w);
        }
      }
    }
  }
  if (LET) {
    __syncthreads();
    for (int i = tid; i < a.K; i += 256) {
      if (want_col) atomicAdd(a.dcol + i, reinterpret_cast<float*>(colacc)[i]);
      if (want_sh) atomicAdd(a.dshift + i, reinterpret_cast<float*>(shacc)[i]);
    }
  }
}

// ================================================================== grouped kernels
template <int LPG>
__device__ __forceinline__ float grp_sum(float v) {
#pragma unroll
  for (int o = LPG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int LPG>
__device__ __forceinline__ float grp_max(float v) {
#pragma unroll
  for (int o = LPG / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// group = 4*LPG*VPL contiguous weights; requires K % group == 0.  LET: col/shift read through L1.
template <int LPG, int VPL, bool LET>
__global__ void __launch_bounds__(256) wq_grp_fwd(WqArgs a) {
  constexpr int GPS = 32 / LPG;        // groups per warp slot
  constexpr int GSZ = 4 * LPG * VPL;
  const int lane = threadIdx.x & 31, sub = lane / LPG, sl_ = lane % LPG;
  const int G = a.K / GSZ;
  const int n_groups = a.N * G;                 // < 2^31 (checked by the dispatcher): 32-bit index arithmetic
  const int n_slots = (n_groups + GPS - 1) / GPS;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nw = gridDim.x * (blockDim.x >> 5);
  const int lwc = a.up != nullptr;
  const bool has_col = LET && a.col != nullptr, has_shift = LET && a.shift != nullptr;
  for (int slot = wid; slot < n_slots; slot += nw) {
    const int item = slot * GPS + sub;
    const bool valid = item < n_groups;
    const int it = valid ? item : n_groups - 1;
    const int n = it / G, gi = it - n * G;
    const long base = (long)it * GSZ;            // groups tile the rows exactly (K % group == 0)
    const int kb = gi * GSZ;
    float4 w[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) w[v] = ld4(a.W + base + 4 * (sl_ + v * LPG));
    const float rs = (LET && a.row) ? a.row[n] : 1.0f;
    const int rop = (LET && a.row) ? a.row_op : 0;
    float mn = INFINITY, mx = -INFINITY, mv = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int k = kb + 4 * (sl_ + v * LPG);
      float4 c4 = has_col ? ld4(a.col + k) : make_float4(1.f, 1.f, 1.f, 1.f);
      if (has_shift) {
        const float4 s4 = ld4(a.shift + k);
        mv += w[v].x * s4.x + w[v].y * s4.y + w[v].z * s4.z + w[v].w * s4.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float f = oq_fold(f4(w[v], j), f4(c4, j), has_col, rs, rop);
        f4(w[v], j) = f;
        mn = fminf(mn, f);
        mx = fmaxf(mx, f);
      }
    }
    mn = -grp_max<LPG>(-mn);
    mx = grp_max<LPG>(mx);
    if (has_shift) mv = grp_sum<LPG>(mv);
    const float su = lwc ? oq_sigmoidf(a.up[it]) : 1.0f;
    const float sl = lwc ? oq_sigmoidf(a.low[it]) : 1.0f;
    const OqQParams q = oq_qparams(mn, mx, su, sl, lwc, a.n_bits, a.symmetric, a.recip_div);
    if (sl_ == 0 && valid) {
      if (a.scale_out) a.scale_out[item] = q.scale;
      if (a.zp_out) a.zp_out[item] = q.zp;
      if (a.xmin_out) a.xmin_out[item] = mn;
      if (a.xmax_out) a.xmax_out[item] = mx;
      if (has_shift && a.mv) atomicAdd(a.mv + n, mv);
    }
    const float inv = OQ_DIV(1.0f, q.scale);
    const int inv_ok = oq_inv_ok(inv);
    if (valid) {
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int ko = 4 * (sl_ + v * LPG);
        float y[4], c[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float t = OQ_ADD(oq_rint_div(f4(w[v], j), q.scale, inv, inv_ok), q.zp);
          c[j] = oq_clampf(t, 0.0f, q.qmax);
          y[j] = oq_dequant(c[j], q);
        }
        if (a.q_bf) store_bf4(a.q_bf + (long)n * a.ld_bf + kb + ko, y);
        if (a.q_f32) *reinterpret_cast<float4*>(a.q_f32 + base + ko) = make_float4(y[0], y[1], y[2], y[3]);
        if (a.codes)
          *reinterpret_cast<uchar4*>(a.codes + base + ko) =
              make_uchar4((unsigned char)c[0], (unsigned char)c[1], (unsigned char)c[2], (unsigned char)c[3]);
      }
    }
  }
}

// dynamic smem (LET): colacc[K] | shacc[K] (shared-memory atomics: several warps hit the same columns)
// TIE: the arg-max / arg-min bookkeeping (who receives the clip-path gradient) is only needed when a gradient w.r.t. the
// weights or the LET vectors is produced; the LWC-only step needs just R per group.
template <int LPG, int VPL, bool LET, bool TIE>
__global__ void __launch_bounds__(256) wq_grp_bwd(WqArgs a) {
  extern __shared__ float4 dsm4[];
  constexpr int GPS = 32 / LPG;
  constexpr int GSZ = 4 * LPG * VPL;
  float* colacc = reinterpret_cast<float*>(dsm4);
  float* shacc = colacc + a.K;
  const bool has_col = LET && a.col != nullptr;
  const bool want_col = LET && a.dcol != nullptr, want_sh = LET && a.dshift != nullptr;
  if (LET) {
    for (int i = threadIdx.x; i < a.K; i += blockDim.x) { colacc[i] = 0.f; shacc[i] = 0.f; }
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, sub = lane / LPG, sl_ = lane % LPG;
  const int G = a.K / GSZ;
  const int n_groups = a.N * G;                 // < 2^31 (checked by the dispatcher): 32-bit index arithmetic
  const int n_slots = (n_groups + GPS - 1) / GPS;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nw = gridDim.x * (blockDim.x >> 5);
  const int lwc = a.up != nullptr;
  for (int slot = wid; slot < n_slots; slot += nw) {
    const int item = slot * GPS + sub;
    const bool valid = item < n_groups;
    const int it = valid ? item : n_groups - 1;
    const int n = it / G, gi = it - n * G;
    const long base = (long)it * GSZ;            // groups tile the rows exactly (K % group == 0)
    const int kb = gi * GSZ;
    float4 w[VPL], g[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int ko = 4 * (sl_ + v * LPG);
      w[v] = ld4(a.W + base + ko);
      if (a.g_is_f32) {
        g[v] = ld4(reinterpret_cast<const float*>(a.g) + (long)n * a.ld_g + kb + ko);
      } else {
        float t[4];
        load_bf4(reinterpret_cast<const bf16*>(a.g) + (long)n * a.ld_g + kb + ko, t);
        g[v] = make_float4(t[0], t[1], t[2], t[3]);
      }
    }
    const float rs = (LET && a.row) ? a.row[n] : 1.0f;
    const int rop = (LET && a.row) ? a.row_op : 0;
    const float rfac = rop == 1 ? rs : (rop == 2 ? 1.0f / rs : 1.0f);
    const float en = (LET && a.e_row) ? a.e_row[n] : 0.0f;
    const float xmin = a.xmin_in[it], xmax = a.xmax_in[it];
    const float su = lwc ? oq_sigmoidf(a.up[it]) : 1.0f;
    const float sl = lwc ? oq_sigmoidf(a.low[it]) : 1.0f;
    OqQParams q;
    if (a.scale_in) {
      q.scale = a.scale_in[it];
      q.zp = a.zp_in[it];
      q.qmax = (float)((1 << a.n_bits) - 1);
    } else {
      q = oq_qparams(xmin, xmax, su, sl, lwc, a.n_bits, a.symmetric, a.recip_div);
    }
    const float inv = OQ_DIV(1.0f, q.scale);
    const int inv_ok = oq_inv_ok(inv);
    float R = 0.f, nmax = 0.f, nmin = 0.f, rowdot = 0.f, smax = 0.f, smin = 0.f;
    float4 fv[VPL];
    const float rmul = rop == 1 ? rs : (rop == 2 ? 1.0f / rs : 1.0f);
    const float tie_tol_max = 2.4e-7f * fabsf(xmax), tie_tol_min = 2.4e-7f * fabsf(xmin);
    const bool tie_all = !(lwc && xmax > 0.0f && xmin < 0.0f) || a.symmetric;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int k = kb + 4 * (sl_ + v * LPG);
      float4 c4 = has_col ? ld4(a.col + k) : make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float wv = f4(w[v], j);
        const float w1 = has_col ? wv * f4(c4, j) : wv;
        const float f = w1 * rmul;
        if (TIE) f4(fv[v], j) = f;
        const float qa = f * inv;
        const float t = rintf(qa) + q.zp;
        const bool inr = (t >= 0.0f) && (t <= q.qmax);
        const float c = oq_clampf(t, 0.0f, q.qmax);
        const float gk = valid ? f4(g[v], j) : 0.0f;
        R += gk * ((c - q.zp) - (inr ? qa : 0.0f));
        if (TIE && (tie_all || t >= q.qmax - 1.0f || t <= 1.0f)) {
          if (fabsf(f - xmax) <= tie_tol_max) { nmax += 1.0f; smax += w1; }
          if (fabsf(f - xmin) <= tie_tol_min) { nmin += 1.0f; smin += w1; }
        }
        const float d2 = inr ? gk : 0.0f;
        if (LET) rowdot += d2 * w1;
        if (TIE) f4(g[v], j) = d2;
        if (valid && want_col) atomicAdd(colacc + k + j, d2 * rfac * wv);
        if (valid && want_sh) atomicAdd(shacc + k + j, wv * en);
      }
    }
    R = grp_sum<LPG>(R);
    if (TIE) { nmax = grp_sum<LPG>(nmax); nmin = grp_sum<LPG>(nmin); }
    float gmaxp, gminp;
    oq_scale_grad(R, xmin, xmax, su, sl, lwc, a.n_bits, a.symmetric, a.recip_div, &gmaxp, &gminp);
    const float cmax = nmax > 0.f ? gmaxp * su / nmax : 0.0f;
    const float cmin = nmin > 0.f ? gminp * sl / nmin : 0.0f;
    if (LET && a.drow) {
      rowdot = grp_sum<LPG>(rowdot); smax = grp_sum<LPG>(smax); smin = grp_sum<LPG>(smin);
      if (sl_ == 0 && valid) atomicAdd(a.drow + n, rowdot + cmax * smax + cmin * smin);
    }
    if (sl_ == 0 && valid && lwc) {
      if (a.d_up) a.d_up[item] = gmaxp * xmax * su * (1.0f - su);
      if (a.d_low) a.d_low[item] = gminp * xmin * sl * (1.0f - sl);
    }
    if (TIE && valid && (want_col || a.dW)) {
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int ko = 4 * (sl_ + v * LPG);
        float4 o4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float f = f4(fv[v], j);
          float corr = 0.0f;
          if (fabsf(f - xmax) <= tie_tol_max) corr += cmax;
          if (fabsf(f - xmin) <= tie_tol_min) corr += cmin;
          if (corr != 0.0f && want_col) atomicAdd(colacc + kb + ko + j, corr * rfac * f4(w[v], j));
          f4(o4, j) = f4(g[v], j) + corr;
        }
        if (a.dW) *reinterpret_cast<float4*>(a.dW + base + ko) = o4;
      }
    }
  }
  if (LET) {
    __syncthreads();
    for (int i = threadIdx.x; i < a.K; i += blockDim.x) {
      if (want_col && colacc[i] != 0.f) atomicAdd(a.dcol + i, colacc[i]);
      if (want_sh && shacc[i] != 0.f) atomicAdd(a.dshift + i, shacc[i]);
    }
  }
}

// ------------------------------------------------------------------ dispatch helpers
template <bool LET, int ROWDIV>
static bool launch_pc_fwd2(const WqArgs& a, int grid, size_t smem, cudaStream_t s) {
  const int nv = (a.K / 4 + 255) / 256;
#define OQ_PC_FWD(NVV)                                                                                          \
  {                                                                                                             \
    if (smem > 40 * 1024) cudaFuncSetAttribute(wq_pc_fwd<NVV, LET, ROWDIV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    wq_pc_fwd<NVV, LET, ROWDIV><<<grid, 256, smem, s>>>(a);                                                     \
    return true;                                                                                                \
  }
  if (nv <= 1) OQ_PC_FWD(1)
  if (nv <= 2) OQ_PC_FWD(2)
  if (nv <= 4) OQ_PC_FWD(4)
  if (nv <= 8) OQ_PC_FWD(8)
  if (nv <= 11) OQ_PC_FWD(11)
  if (nv <= 16) OQ_PC_FWD(16)
#undef OQ_PC_FWD
  return false;
}
// calibration-loop variant: bf16 output (+ s8 side output) only
template <int NV, int THREADS, bool LET, int ROWMODE, bool S8, bool EXACT>
static bool launch_pc_fwd_lean4(const WqArgs& a, size_t smem, cudaStream_t s) {
  auto kern = wq_pc_fwd_lean<NV, THREADS, LET, ROWMODE, S8, EXACT>;
  static int per_sm = 0;
  if (smem > 96 * 1024) return false;
  if (!per_sm) {
    if (smem > 40 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess) return false;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem) != cudaSuccess || per_sm < 1) {
      per_sm = 0;
      cudaGetLastError();
      return false;
    }
  }
  const int grid = (int)std::min<long>(a.N, (long)oq_num_sms() * per_sm);
  kern<<<grid, THREADS, smem, s>>>(a);
  return true;
}
template <int NV, int THREADS, bool LET, int ROWMODE, bool EXACT>
static bool launch_pc_fwd_lean3(const WqArgs& a, size_t smem, cudaStream_t s) {
  return a.codes8 ? launch_pc_fwd_lean4<NV, THREADS, LET, ROWMODE, true, EXACT>(a, smem, s)
                  : launch_pc_fwd_lean4<NV, THREADS, LET, ROWMODE, false, EXACT>(a, smem, s);
}
// Instantiations are kept few (compile time): 128 threads for K <= 4096, 256 beyond; exact row fills for the shapes of
// the model families, a bounds-checked variant for everything else; row folds (none / multiply / divide / segment table)
// all go through the per-row table variant (ROWMODE 2), only "no row fold at all" has its own.
template <bool LET, int ROWMODE>
static bool launch_pc_fwd_lean2(const WqArgs& a, size_t smem, cudaStream_t s) {
  const int K4 = a.K / 4;
  if (K4 <= 1024) {
    const int nv = (K4 + 127) / 128;
    if (nv == 2) return launch_pc_fwd_lean3<2, 128, LET, ROWMODE, true>(a, smem, s);
    if (nv == 4) return launch_pc_fwd_lean3<4, 128, LET, ROWMODE, true>(a, smem, s);
    if (nv == 6) return launch_pc_fwd_lean3<6, 128, LET, ROWMODE, true>(a, smem, s);
    if (nv == 8) return launch_pc_fwd_lean3<8, 128, LET, ROWMODE, true>(a, smem, s);
    return launch_pc_fwd_lean3<8, 128, LET, ROWMODE, false>(a, smem, s);
  }
  const int nv = (K4 + 255) / 256;
  if (nv == 8) return launch_pc_fwd_lean3<8, 256, LET, ROWMODE, true>(a, smem, s);
  if (nv == 11) return launch_pc_fwd_lean3<11, 256, LET, ROWMODE, true>(a, smem, s);
  if (nv == 16) return launch_pc_fwd_lean3<16, 256, LET, ROWMODE, true>(a, smem, s);
  if (nv <= 8) return launch_pc_fwd_lean3<8, 256, LET, ROWMODE, false>(a, smem, s);
  if (nv <= 16) return launch_pc_fwd_lean3<16, 256, LET, ROWMODE, false>(a, smem, s);
  return false;
}
static bool launch_pc_fwd_lean(const WqArgs& a, bool let, size_t smem, cudaStream_t s) {
  static const bool off = [] { const char* e = getenv("OQ_WQ_LEAN"); return e && atoi(e) == 0; }();
  if (off || !a.q_bf || a.q_f32 || a.codes) return false;
  if (!let) return launch_pc_fwd_lean2<false, 3>(a, 0, s);
  if (a.nseg || (a.row && a.row_op != 0)) return launch_pc_fwd_lean2<true, 2>(a, smem, s);
  return launch_pc_fwd_lean2<true, 3>(a, smem, s);
}
template <bool LET>
static bool launch_pc_fwd(const WqArgs& a, int grid, size_t smem, cudaStream_t s) {
  if (launch_pc_fwd_lean(a, LET, smem, s)) return true;
  if (LET && a.nseg) return launch_pc_fwd2<LET, 2>(a, grid, smem, s);
  if (LET && a.row && a.row_op == 2) return launch_pc_fwd2<LET, 1>(a, grid, smem, s);
  return launch_pc_fwd2<LET, 0>(a, grid, smem, s);
}
template <bool LET>
static bool launch_pc_bwd(const WqArgs& a, int grid, size_t smem, cudaStream_t s) {
  const int nv = (a.K / 4 + 255) / 256;
#define OQ_PC_BWD(NVV)                                                                                          \
  {                                                                                                             \
    if (smem > 40 * 1024) cudaFuncSetAttribute(wq_pc_bwd<NVV, LET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    wq_pc_bwd<NVV, LET><<<grid, 256, smem, s>>>(a);                                                             \
    return true;                                                                                                \
  }
  if (nv <= 1) OQ_PC_BWD(1)
  if (nv <= 2) OQ_PC_BWD(2)
  if (nv <= 4) OQ_PC_BWD(4)
  if (nv <= 8) OQ_PC_BWD(8)
  if (nv <= 11) OQ_PC_BWD(11)
  if (nv <= 16) OQ_PC_BWD(16)
#undef OQ_PC_BWD
  return false;
}
// Elements per lane of the grouped kernels.  The per-group scalar work (two sigmoids, the scale / zero-point /
// reciprocal divisions, the shuffles) is executed by every lane of a group, so few lanes per group amortise it best:
// measured on B200 (tools/wq_bench.py, g128) forward 2.0 / 3.9 / 4.6 TB/s and backward 1.27 / 2.4 / 2.4 TB/s at
// 4 / 16 / 32 elements per lane.  OQ_WQ_EPL_FWD / OQ_WQ_EPL_BWD override for experiments.
static int grp_epl(int group, bool bwd) {
  static int env_f = [] { const char* e = getenv("OQ_WQ_EPL_FWD"); return e ? atoi(e) : 0; }();
  static int env_b = [] { const char* e = getenv("OQ_WQ_EPL_BWD"); return e ? atoi(e) : 0; }();
  int epl = bwd ? (env_b ? env_b : 16) : (env_f ? env_f : 32);
  if (group == 512 && epl == 4) epl = 16;
  if (group == 256 && epl == 4) epl = 8;
  if (epl > group) epl = group;
  return epl;
}
static int grp_lpg(int group, bool bwd) { return group / grp_epl(group, bwd); }

template <bool LET, bool BWD>
static bool launch_grp(const WqArgs& a, int grid, size_t smem, cudaStream_t s) {
  const int epl = grp_epl(a.group, BWD);
#define OQ_GRP(L, V)                                                                                             \
  {                                                                                                              \
    if (BWD) {                                                                                                   \
      if (LET || a.dW) {                                                                                         \
        if (smem > 40 * 1024) cudaFuncSetAttribute(wq_grp_bwd<L, V, LET, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        wq_grp_bwd<L, V, LET, true><<<grid, 256, smem, s>>>(a);                                                  \
      } else {                                                                                                   \
        wq_grp_bwd<L, V, false, false><<<grid, 256, 0, s>>>(a);                                                  \
      }                                                                                                          \
    } else {                                                                                                     \
      wq_grp_fwd<L, V, LET><<<grid, 256, 0, s>>>(a);                                                             \
    }                                                                                                            \
    return true;                                                                                                 \
  }
  // elements per lane: 16 (4 float4) -- the per-group scalar work (two sigmoids, scale / zero-point / reciprocal
  // divisions, the shuffles) is executed by every lane of a group, so few lanes per group amortise it best
  switch (a.group * 100 + epl) {
    case 3204: OQ_GRP(8, 1)
    case 6404: OQ_GRP(16, 1)
    case 12804: OQ_GRP(32, 1)
    case 25608: OQ_GRP(32, 2)
    case 3216: OQ_GRP(2, 4)
    case 6416: OQ_GRP(4, 4)
    case 12816: OQ_GRP(8, 4)
    case 25616: OQ_GRP(16, 4)
    case 51216: OQ_GRP(32, 4)
    case 3232: OQ_GRP(1, 8)
    case 6432: OQ_GRP(2, 8)
    case 12832: OQ_GRP(4, 8)
    case 25632: OQ_GRP(8, 8)
    case 51232: OQ_GRP(16, 8)
    default: return false;
  }
#undef OQ_GRP
}
