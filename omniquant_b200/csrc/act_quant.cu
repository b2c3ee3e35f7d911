// K4: per-token dynamic activation fake-quant (forward + backward), bf16 in/out,
// fp32 arithmetic.  Optional fused rotary embedding for the Q/K operands of
// QuantMatMul.
//
// Reference: UniformAffineQuantizer with dynamic_method="per_token"
// (quantize/quantizer.py:108-147, :84-105) as called from QuantLinear.forward
// (quantize/int_linear.py:59-60) and QuantMatMul.quant_x1/x2
// (quantize/int_matmul.py:31-39; LLaMA call sites models/int_llama_layer.py:124-125,
// :141-142,:162).  Backward = autograd of the same: straight
// through inside the clamp, plus the scale path into the arg-max / arg-min
// elements (split evenly over ties); nothing through the rounded zero point.
//
// HBM-bound: 2 B read + 2 B write per activation.  Rows are addressed with two
// levels (outer token stride, inner head stride) so the [T, 3H] output of the
// fused QKV GEMM is consumed in place.
#include <algorithm>

#include "common.cuh"

namespace {

struct RowIx {
  long rows;
  int rows_inner;           // rows per outer index (heads per token), 1 for plain 2-D
  long ld_outer, ld_inner;  // element strides
  __device__ __forceinline__ long off(long r) const {
    if (rows_inner == 1) return r * ld_outer;
    const unsigned o = (unsigned)r / (unsigned)rows_inner;   // rows < 2^31 is checked at the entry points
    return (long)o * ld_outer + (long)((unsigned)r - o * (unsigned)rows_inner) * ld_inner;
  }
};

struct AqArgs {
  const bf16* x; RowIx xi;
  bf16* y; RowIx yi;
  const bf16* gy; RowIx gi;
  bf16* gx; RowIx gxi;
  int C, n_bits, symmetric, recip_div;
  const float* cos; const float* sin; int seq_len;  // rope tables [seq_len, C] or null
  int rope_inner;  // inner rows [0, rope_inner) are rotated
  float* scale_out; float* zp_out;
  S8Side s8;   // optional exact-integer side output (plain 2-D rows only)
};

// ---- element arithmetic shared by the register-resident kernels ---------------------------------------------------
// t = rint(x / scale) + zp.  FAST: the quotient is x*inv corrected by one fused residual step,
// q1 = fma(fma(-q0, scale, x), inv, q0) with inv = fl(1/scale) -- Markstein's correction, which returns the IEEE
// quotient fl(x/scale) (exhaustively compared with the true division on 2e9 random pairs, tools/div_check.c) at 3
// instructions instead of the ~20 of an IEEE division.  !FAST (scale 0 / denormal / huge: inv not a normal number):
// the IEEE division itself, so a constant row behaves exactly like the reference (NaN / saturated codes).
template <bool FAST>
__device__ __forceinline__ float aq_t(float x, const OqQParams& q, float inv, float& qa) {
  qa = __fmul_rn(x, inv);
  if constexpr (FAST) return OQ_ADD(rintf(__fmaf_rn(__fmaf_rn(-qa, q.scale, x), inv, qa)), q.zp);
  else return OQ_ADD(rintf(OQ_DIV(x, q.scale)), q.zp);
}
template <bool FAST, int E>
__device__ __forceinline__ void aq_quant_elems_t(float (&v)[E], const OqQParams& q, float inv) {
#pragma unroll
  for (int j = 0; j < E; ++j) {
    float qa;
    v[j] = oq_dequant(oq_clampf(aq_t<FAST>(v[j], q, inv, qa), 0.0f, q.qmax), q);
  }
}
template <int E>
__device__ __forceinline__ void aq_quant_elems(float (&v)[E], const OqQParams& q, float inv, int inv_ok) {
  if (inv_ok) aq_quant_elems_t<true, E>(v, q, inv);
  else aq_quant_elems_t<false, E>(v, q, inv);
}
// same, also returning the clamped integer codes (for the s8 side output)
template <bool FAST, int E>
__device__ __forceinline__ void aq_quant_codes_t(float (&v)[E], float (&c)[E], const OqQParams& q, float inv) {
#pragma unroll
  for (int j = 0; j < E; ++j) {
    float qa;
    c[j] = oq_clampf(aq_t<FAST>(v[j], q, inv, qa), 0.0f, q.qmax);
    v[j] = oq_dequant(c[j], q);
  }
}
template <int E>
__device__ __forceinline__ void aq_quant_codes(float (&v)[E], float (&c)[E], const OqQParams& q, float inv, int inv_ok) {
  if (inv_ok) aq_quant_codes_t<true, E>(v, c, q, inv);
  else aq_quant_codes_t<false, E>(v, c, q, inv);
}
// first backward pass over E elements: accumulates R = dL/dscale, the tie counts of the row extrema, and zeroes the
// gradient of clipped elements (straight-through part); same codes as the forward, ties included.
// The in-range test and the clamp are one operation: with t = rint(q) + zp and zp an integer, t in [0, qmax] <=>
// qlo <= q <= qhi for the half-integers qlo = -zp - 0.5, qhi = qmax - zp + 0.5 (moved one ulp inwards when the tie
// rounds -- to even -- out of the range), and rint(clamp(q, qlo, qhi)) = clamp(t, 0, qmax) - zp; rint of the bounded value
// is a magic-number add.  The extrema are counted only by the threads whose 8..16 elements contain one (a 3-input max /
// min per element pair instead of two compare-and-adds per element).  Identical values for finite inputs.
struct AqRange { float qlo, qhi; };
__device__ __forceinline__ AqRange aq_range(const OqQParams& q) {
  const float lo = -q.zp, hi = q.qmax - q.zp;                   // integers (|zp| <= 1e4): exact
  const bool lo_even = rintf(lo * 0.5f) * 2.0f == lo, hi_even = rintf(hi * 0.5f) * 2.0f == hi;
  float qlo = lo - 0.5f, qhi = hi + 0.5f;                        // never zero
  if (!lo_even) qlo = __int_as_float(__float_as_int(qlo) + (qlo > 0.f ? 1 : -1));     // one ulp up
  if (!hi_even) qhi = __int_as_float(__float_as_int(qhi) + (qhi > 0.f ? -1 : 1));     // one ulp down
  return {qlo, qhi};
}
// returns whether one of the E elements is a row extremum (only those need the tie correction of the last pass)
template <bool FAST, int E>
__device__ __forceinline__ bool aq_bwd_elems_t(const float (&v)[E], float (&g)[E], const OqQParams& q, float inv, AqRange rg,
                                               float mn, float mx, float& R, float& nmax, float& nmin) {
  const float magic = 12582912.0f;
  float vhi = -INFINITY, vlo = INFINITY;
#pragma unroll
  for (int j = 0; j < E; ++j) {
    const float qa = __fmul_rn(v[j], inv);
    const float q1 = FAST ? __fmaf_rn(__fmaf_rn(-qa, q.scale, v[j]), inv, qa) : OQ_DIV(v[j], q.scale);
    const float qc = fminf(fmaxf(q1, rg.qlo), rg.qhi);
    const bool inr = qc == q1;
    const float rc = __fsub_rn(__fadd_rn(qc, magic), magic);       // clamp(rint(q1) + zp, 0, qmax) - zp
    R = fmaf(g[j], rc - (inr ? qa : 0.0f), R);
    g[j] = inr ? g[j] : 0.0f;
    vhi = fmaxf(vhi, v[j]);
    vlo = fminf(vlo, v[j]);
  }
  const bool ext = vhi == mx || vlo == mn;
  if (ext) {
#pragma unroll
    for (int j = 0; j < E; ++j) {
      nmax += (v[j] == mx) ? 1.0f : 0.0f;
      nmin += (v[j] == mn) ? 1.0f : 0.0f;
    }
  }
  return ext;
}
template <int E>
__device__ __forceinline__ bool aq_bwd_elems(const float (&v)[E], float (&g)[E], const OqQParams& q, float inv, int inv_ok,
                                             AqRange rg, float mn, float mx, float& R, float& nmax, float& nmin) {
  if (inv_ok) return aq_bwd_elems_t<true, E>(v, g, q, inv, rg, mn, mx, R, nmax, nmin);
  return aq_bwd_elems_t<false, E>(v, g, q, inv, rg, mn, mx, R, nmax, nmin);
}

// ---- block per row (C large, no rope) ---------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256) aq_fwd_block(AqArgs a) {
  __shared__ float scratch[2 * 32];
  const long r = blockIdx.x;
  const bf16* xr = a.x + a.xi.off(r);
  bf16* yr = a.y + a.yi.off(r);
  constexpr int V = VEC ? 8 : 1;
  float mm[2] = {-INFINITY, -INFINITY};  // -min, max
  for (int c = threadIdx.x * V; c < a.C; c += blockDim.x * V) {
    float v[V];
    if constexpr (VEC) load_bf8(xr + c, v);
    else v[0] = bf2f(xr[c]);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      mm[0] = fmaxf(mm[0], -v[j]);
      mm[1] = fmaxf(mm[1], v[j]);
    }
  }
  block_reduce<2>(mm, scratch, OpMax(), -INFINITY);
  const OqQParams q = oq_qparams(-mm[0], mm[1], 1.f, 1.f, 0, a.n_bits, a.symmetric, a.recip_div);
  if (threadIdx.x == 0) {
    if (a.scale_out) a.scale_out[r] = q.scale;
    if (a.zp_out) a.zp_out[r] = q.zp;
  }
  for (int c = threadIdx.x * V; c < a.C; c += blockDim.x * V) {
    float v[V];
    if constexpr (VEC) load_bf8(xr + c, v);
    else v[0] = bf2f(xr[c]);
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = oq_fake_quant(v[j], q);
    if constexpr (VEC) store_bf8(yr + c, v);
    else yr[c] = f2bf(v[0]);
  }
}

template <bool VEC>
__global__ void __launch_bounds__(256) aq_bwd_block(AqArgs a) {
  __shared__ float scratch[3 * 32];
  const long r = blockIdx.x;
  const bf16* xr = a.x + a.xi.off(r);
  const bf16* gr = a.gy + a.gi.off(r);
  bf16* dr = a.gx + a.gxi.off(r);
  constexpr int V = VEC ? 8 : 1;
  float mm[2] = {-INFINITY, -INFINITY};
  for (int c = threadIdx.x * V; c < a.C; c += blockDim.x * V) {
    float v[V];
    if constexpr (VEC) load_bf8(xr + c, v);
    else v[0] = bf2f(xr[c]);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      mm[0] = fmaxf(mm[0], -v[j]);
      mm[1] = fmaxf(mm[1], v[j]);
    }
  }
  block_reduce<2>(mm, scratch, OpMax(), -INFINITY);
  const float xmin = -mm[0], xmax = mm[1];
  const OqQParams q = oq_qparams(xmin, xmax, 1.f, 1.f, 0, a.n_bits, a.symmetric, a.recip_div);
  float acc[3] = {0.f, 0.f, 0.f};  // R, #max, #min
  for (int c = threadIdx.x * V; c < a.C; c += blockDim.x * V) {
    float v[V], g[V];
    if constexpr (VEC) { load_bf8(xr + c, v); load_bf8(gr + c, g); }
    else { v[0] = bf2f(xr[c]); g[0] = bf2f(gr[c]); }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float t = oq_code_unclamped(v[j], q);
      const bool inr = (t >= 0.0f) && (t <= q.qmax);
      const float cc = oq_clampf(t, 0.0f, q.qmax);
      acc[0] += g[j] * ((cc - q.zp) - (inr ? v[j] / q.scale : 0.0f));
      acc[1] += (v[j] == xmax) ? 1.0f : 0.0f;
      acc[2] += (v[j] == xmin) ? 1.0f : 0.0f;
    }
  }
  block_reduce<3>(acc, scratch, OpSum(), 0.0f);
  float gmaxp, gminp;
  oq_scale_grad(acc[0], xmin, xmax, 1.f, 1.f, 0, a.n_bits, a.symmetric, a.recip_div, &gmaxp, &gminp);
  const float cmax = gmaxp / acc[1], cmin = gminp / acc[2];
  for (int c = threadIdx.x * V; c < a.C; c += blockDim.x * V) {
    float v[V], g[V];
    if constexpr (VEC) { load_bf8(xr + c, v); load_bf8(gr + c, g); }
    else { v[0] = bf2f(xr[c]); g[0] = bf2f(gr[c]); }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float t = oq_code_unclamped(v[j], q);
      const bool inr = (t >= 0.0f) && (t <= q.qmax);
      float d = inr ? g[j] : 0.0f;
      if (v[j] == xmax) d += cmax;
      if (v[j] == xmin) d += cmin;
      g[j] = d;
    }
    if constexpr (VEC) store_bf8(dr + c, g);
    else dr[c] = f2bf(g[0]);
  }
}


// ---- register-resident block-per-row variants (C <= 2048*NV, C % 8 == 0): one read of x (and gy), one write
template <int NV>
__global__ void __launch_bounds__(256) aq_fwd_block_reg(AqArgs a) {
  __shared__ float scratch[3 * 8];
  const long r = blockIdx.x;
  const bf16* xr = a.x + a.xi.off(r);
  bf16* yr = a.y + a.yi.off(r);
  float v[NV][8];
  float mn = INFINITY, mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (threadIdx.x + i * 256) * 8;
    if (c < a.C) {
      load_bf8(xr + c, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) { mn = fminf(mn, v[i][j]); mx = fmaxf(mx, v[i][j]); }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  mn = warp_reduce(mn, OpMin());
  mx = warp_reduce(mx, OpMax());
  if (lane == 0) { scratch[warp] = mn; scratch[8 + warp] = mx; }
  __syncthreads();
  mn = scratch[0]; mx = scratch[8];
#pragma unroll
  for (int w = 1; w < 8; ++w) { mn = fminf(mn, scratch[w]); mx = fmaxf(mx, scratch[8 + w]); }
  const OqQParams q = oq_qparams(mn, mx, 1.f, 1.f, 0, a.n_bits, a.symmetric, a.recip_div);
  if (threadIdx.x == 0) {
    if (a.scale_out) a.scale_out[r] = q.scale;
    if (a.zp_out) a.zp_out[r] = q.zp;
  }
  const float inv = OQ_DIV(1.0f, q.scale);
  const int inv_ok = oq_inv_ok(inv);
  if (!a.s8.codes) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (threadIdx.x + i * 256) * 8;
      if (c < a.C) {
        aq_quant_elems<8>(v[i], q, inv, inv_ok);
        store_bf8(yr + c, v[i]);
      }
    }
    return;
  }
  const float z0 = oq_s8_z0(q.zp, q.qmax);
  int isum = 0;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (threadIdx.x + i * 256) * 8;
    if (c < a.C) {
      float cd[8];
      aq_quant_codes<8>(v[i], cd, q, inv, inv_ok);
      store_bf8(yr + c, v[i]);
      s8_store8(a.s8.codes + r * a.s8.ld + c, cd, z0, isum);
    }
  }
  const float csum = (float)__reduce_add_sync(0xffffffffu, isum);
  if (lane == 0) scratch[16 + warp] = csum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += scratch[16 + w];
    s8_write_meta(a.s8, r, q, z0, t);
  }
}

template <int NV>
__global__ void __launch_bounds__(256) aq_bwd_block_reg(AqArgs a) {
  __shared__ float scratch[5 * 8];
  const long r = blockIdx.x;
  const bf16* xr = a.x + a.xi.off(r);
  const bf16* gr = a.gy + a.gi.off(r);
  bf16* dr = a.gx + a.gxi.off(r);
  float v[NV][8], g[NV][8];
  float mn = INFINITY, mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (threadIdx.x + i * 256) * 8;
    if (c < a.C) {
      load_bf8(xr + c, v[i]);
      load_bf8(gr + c, g[i]);
#pragma unroll
      for (int j = 0; j < 8; j += 2) { mn = fminf(fminf(mn, v[i][j]), v[i][j + 1]); mx = fmaxf(fmaxf(mx, v[i][j]), v[i][j + 1]); }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  mn = warp_reduce(mn, OpMin());
  mx = warp_reduce(mx, OpMax());
  if (lane == 0) { scratch[warp] = mn; scratch[8 + warp] = mx; }
  __syncthreads();
  mn = scratch[0]; mx = scratch[8];
#pragma unroll
  for (int w = 1; w < 8; ++w) { mn = fminf(mn, scratch[w]); mx = fmaxf(mx, scratch[8 + w]); }
  const OqQParams q = oq_qparams(mn, mx, 1.f, 1.f, 0, a.n_bits, a.symmetric, a.recip_div);
  const float inv = OQ_DIV(1.0f, q.scale);
  const int inv_ok = oq_inv_ok(inv);
  const AqRange rg = aq_range(q);
  float R = 0.f, nmax = 0.f, nmin = 0.f;
  bool ext[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (threadIdx.x + i * 256) * 8;
    ext[i] = false;
    if (c < a.C) ext[i] = aq_bwd_elems<8>(v[i], g[i], q, inv, inv_ok, rg, mn, mx, R, nmax, nmin);
  }
  R = warp_reduce(R, OpSum()); nmax = warp_reduce(nmax, OpSum()); nmin = warp_reduce(nmin, OpSum());
  if (lane == 0) { scratch[16 + warp] = R; scratch[24 + warp] = nmax; scratch[32 + warp] = nmin; }
  __syncthreads();
  R = nmax = nmin = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) { R += scratch[16 + w]; nmax += scratch[24 + w]; nmin += scratch[32 + w]; }
  float gmaxp, gminp;
  oq_scale_grad(R, mn, mx, 1.f, 1.f, 0, a.n_bits, a.symmetric, a.recip_div, &gmaxp, &gminp);
  const float cmax = gmaxp / nmax, cmin = gminp / nmin;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (threadIdx.x + i * 256) * 8;
    if (c < a.C) {
      if (ext[i]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (v[i][j] == mx) g[i][j] += cmax;
          if (v[i][j] == mn) g[i][j] += cmin;
        }
      }
      store_bf8(dr + c, g[i]);
    }
  }
}

// ---- SwiGLU fused with the per-token quantiser of down_proj's input (F <= 4096*NV, F % 8 == 0) -----------------
// m = bf16(silu(gate) * up) is never written: the forward stores only fake_quant(m), the backward re-derives m from
// [gate | up] (same expression, same rounding), applies the quantiser backward and the SwiGLU backward in registers.
// 512 threads per token row; the row stays packed (bf16 pairs) in registers between the phases.
template <int NW>
__device__ __forceinline__ void block_minmax(float& mn, float& mx, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  mn = warp_reduce(mn, OpMin());
  mx = warp_reduce(mx, OpMax());
  if (lane == 0) { scratch[warp] = mn; scratch[NW + warp] = mx; }
  __syncthreads();
  mn = scratch[0]; mx = scratch[NW];
#pragma unroll
  for (int w = 1; w < NW; ++w) { mn = fminf(mn, scratch[w]); mx = fmaxf(mx, scratch[NW + w]); }
}
__device__ __forceinline__ void unpack_u4(const uint4& u, float (&v)[8]) {
  unpack_bf2(u.x, v[0], v[1]); unpack_bf2(u.y, v[2], v[3]); unpack_bf2(u.z, v[4], v[5]); unpack_bf2(u.w, v[6], v[7]);
}
__device__ __forceinline__ uint4 pack_u4(const float (&v)[8]) {
  return make_uint4(pack_bf2(v[0], v[1]), pack_bf2(v[2], v[3]), pack_bf2(v[4], v[5]), pack_bf2(v[6], v[7]));
}

struct GluArgs {
  const bf16* gu; long ldgu;     // [rows, 2F] = [gate | up]
  bf16* mq; long ldmq;           // forward output [rows, F]
  const bf16* gy; long ldgy;     // backward: dL/d mq
  bf16* dgu; long lddgu;         // backward output [rows, 2F]
  int F, n_bits, symmetric, recip_div;
  S8Side s8;                     // optional exact-integer side output of the forward
};

template <int NV>
__global__ void __launch_bounds__(512, 2) swiglu_aq_fwd_kernel(GluArgs a) {
  __shared__ float scratch[3 * 16];
  const long r = blockIdx.x;
  const bf16* gr = a.gu + r * a.ldgu;
  uint4 rg[NV], ru[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (threadIdx.x + i * 512) * 8;
    if (c < a.F) {
      rg[i] = *reinterpret_cast<const uint4*>(gr + c);
      ru[i] = *reinterpret_cast<const uint4*>(gr + a.F + c);
    }
  }
  float v[NV][8];
  float mn = INFINITY, mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (threadIdx.x + i * 512) * 8;
    if (c < a.F) {
      float g[8], u[8];
      unpack_u4(rg[i], g);
      unpack_u4(ru[i], u);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[i][j] = bf16_round(swiglu_val(g[j], u[j]));
        mn = fminf(mn, v[i][j]); mx = fmaxf(mx, v[i][j]);
      }
    }
  }
  block_minmax<16>(mn, mx, scratch);
  const OqQParams q = oq_qparams(mn, mx, 1.f, 1.f, 0, a.n_bits, a.symmetric, a.recip_div);
  const float inv = OQ_DIV(1.0f, q.scale);
  const int inv_ok = oq_inv_ok(inv);
  bf16* yr = a.mq + r * a.ldmq;
  if (!a.s8.codes) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (threadIdx.x + i * 512) * 8;
      if (c < a.F) {
        aq_quant_elems<8>(v[i], q, inv, inv_ok);
        store_bf8(yr + c, v[i]);
      }
    }
    return;
  }
  const float z0 = oq_s8_z0(q.zp, q.qmax);
  int isum = 0;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (threadIdx.x + i * 512) * 8;
    if (c < a.F) {
      float cd[8];
      aq_quant_codes<8>(v[i], cd, q, inv, inv_ok);
      store_bf8(yr + c, v[i]);
      s8_store8(a.s8.codes + r * a.s8.ld + c, cd, z0, isum);
    }
  }
  const float csum = (float)__reduce_add_sync(0xffffffffu, isum);
  if ((threadIdx.x & 31) == 0) scratch[32 + (threadIdx.x >> 5)] = csum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 16; ++w) t += scratch[32 + w];
    s8_write_meta(a.s8, r, q, z0, t);
  }
}

// One 512-thread block per row, TWO blocks per SM: the row passes three phases separated by block reductions (extrema of
// m, then R / tie counts), and with every operand in registers (128 per thread, one block per SM) nothing overlapped the
// reductions or the loads (ncu: 93 us for 225 MB).  Each thread now parks its slice of m and of dL/dmq in shared memory
// (the gradient arrives there straight from global memory with cp.async; a thread only ever reads back what it stored,
// so no barrier guards the buffer) and keeps gate / up packed in registers: <= 64 registers, two rows in flight per SM.
template <int NV>
__global__ void __launch_bounds__(512, 2) swiglu_aq_bwd_kernel(GluArgs a) {
  extern __shared__ uint4 park[];            // [2][NV * 512]: m, dL/dmq
  __shared__ float scratch[5 * 16];
  const long r = blockIdx.x;
  const bf16* gr = a.gu + r * a.ldgu;
  const bf16* yr = a.gy + r * a.ldgy;
  uint4* sm_m = park + threadIdx.x;
  uint4* sm_y = park + NV * 512 + threadIdx.x;
  uint4 rg[NV], ru[NV];                      // gate, up: bf16-exact, kept packed
  float mn = INFINITY, mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (threadIdx.x + i * 512) * 8;
    if (c < a.F) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sm_y + i * 512)), "l"(yr + c) : "memory");
      rg[i] = *reinterpret_cast<const uint4*>(gr + c);
      ru[i] = *reinterpret_cast<const uint4*>(gr + a.F + c);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (threadIdx.x + i * 512) * 8;
    if (c < a.F) {
      float g[8], u[8];
      unpack_u4(rg[i], g);
      unpack_u4(ru[i], u);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = swiglu_val(g[j], u[j]);
      const uint4 m = pack_u4(g);
      sm_m[i * 512] = m;
      unpack_u4(m, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) { mn = fminf(mn, g[j]); mx = fmaxf(mx, g[j]); }
    }
  }
  block_minmax<16>(mn, mx, scratch);
  const OqQParams q = oq_qparams(mn, mx, 1.f, 1.f, 0, a.n_bits, a.symmetric, a.recip_div);
  const float inv = OQ_DIV(1.0f, q.scale);
  const int inv_ok = oq_inv_ok(inv);
  const AqRange qr = aq_range(q);
  float R = 0.f, nmax = 0.f, nmin = 0.f;
  bool ext[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) ext[i] = false;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (threadIdx.x + i * 512) * 8;
    if (c < a.F) {
      float x[8], gy[8];
      unpack_u4(sm_m[i * 512], x);
      unpack_u4(sm_y[i * 512], gy);
      ext[i] = aq_bwd_elems<8>(x, gy, q, inv, inv_ok, qr, mn, mx, R, nmax, nmin);
      sm_y[i * 512] = pack_u4(gy);       // exact: entries are bf16 values or zero
    }
  }
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    R = warp_reduce(R, OpSum()); nmax = warp_reduce(nmax, OpSum()); nmin = warp_reduce(nmin, OpSum());
    if (lane == 0) { scratch[32 + warp] = R; scratch[48 + warp] = nmax; scratch[64 + warp] = nmin; }
    __syncthreads();
    R = nmax = nmin = 0.f;
#pragma unroll
    for (int w = 0; w < 16; ++w) { R += scratch[32 + w]; nmax += scratch[48 + w]; nmin += scratch[64 + w]; }
  }
  float gmaxp, gminp;
  oq_scale_grad(R, mn, mx, 1.f, 1.f, 0, a.n_bits, a.symmetric, a.recip_div, &gmaxp, &gminp);
  const float cmax = gmaxp / nmax, cmin = gminp / nmin;
  bf16* dr = a.dgu + r * a.lddgu;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (threadIdx.x + i * 512) * 8;
    if (c < a.F) {
      float d[8], g[8], u[8];
      unpack_u4(sm_y[i * 512], d);
      unpack_u4(rg[i], g);
      unpack_u4(ru[i], u);
      if (ext[i]) {
        float x[8];
        unpack_u4(sm_m[i * 512], x);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (x[j] == mx) d[j] += cmax;
          if (x[j] == mn) d[j] += cmin;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float dg, du;
        swiglu_grad(g[j], u[j], bf16_round(d[j]), dg, du);   // dL/dm passes through bf16 like the unfused path
        g[j] = dg;
        u[j] = du;
      }
      store_bf8(dr + c, g);
      store_bf8(dr + a.F + c, u);
    }
  }
}

// ---- eight lanes per row (C = 64 or 128, optional rope) -----------------------
// Eight lanes own one row (head) of C = 16*HALF elements: lane s holds [HALF*s, HALF*s+HALF) and the same range of
// the second half, so the rotary partner x[i +- C/2] sits in the same lane (no shuffle) and the row reductions take
// three shuffle steps.  A warp works on four rows at a time and on AQ_HPW consecutive heads of one token: the rotary
// tables of the token are read once into registers; all index arithmetic is 32-bit and per warp.
// Inner rows >= a.rope_inner are not rotated (V in the fused [q | k | v] buffer).
constexpr int AQ_HPW = 8;

template <typename Op>
__device__ __forceinline__ float group8_reduce(float v, Op op) {
  v = op(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = op(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = op(v, __shfl_xor_sync(0xffffffffu, v, 4));
  return v;
}

template <int HALF>
__device__ __forceinline__ void load_half(const bf16* p, float* v) {
  if constexpr (HALF == 8) { float t[8]; load_bf8(p, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = t[j];
  } else { float t[4]; load_bf4(p, t);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = t[j];
  }
}
template <int HALF>
__device__ __forceinline__ void store_half(bf16* p, const float* v) {
  if constexpr (HALF == 8) { float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = v[j];
    store_bf8(p, t);
  } else { float t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = v[j];
    store_bf4(p, t);
  }
}

// y = x*cos + rotate_half(x)*sin ; rotate_half(x)[i] = -x[i+h] (i<h), x[i-h] (i>=h)
// (transformers 4.31 apply_rotary_pos_emb as used at models/int_llama_layer.py:125); v = [first half | second half]
template <int HALF>
__device__ __forceinline__ void rope_fwd(float (&v)[2 * HALF], const float (&cs)[2 * HALF], const float (&sn)[2 * HALF]) {
#pragma unroll
  for (int j = 0; j < HALF; ++j) {
    const float lo = v[j], hi = v[HALF + j];
    v[j] = lo * cs[j] + (-hi) * sn[j];
    v[HALF + j] = hi * cs[HALF + j] + lo * sn[HALF + j];
  }
}
// transpose of the above: dx[j] = g[j]cos[j] + (j<h ? g[j+h]sin[j+h] : -g[j-h]sin[j-h])
template <int HALF>
__device__ __forceinline__ void rope_bwd(float (&g)[2 * HALF], const float (&cs)[2 * HALF], const float (&sn)[2 * HALF]) {
#pragma unroll
  for (int j = 0; j < HALF; ++j) {
    const float lo = g[j], hi = g[HALF + j];
    g[j] = lo * cs[j] + hi * sn[HALF + j];
    g[HALF + j] = hi * cs[HALF + j] - lo * sn[j];
  }
}

template <int E>
__device__ __forceinline__ void aq_row_fwd(float (&v)[E], const AqArgs& a, long r, bool writer) {
  float mn = INFINITY, mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < E; ++j) { mn = fminf(mn, v[j]); mx = fmaxf(mx, v[j]); }
  mn = group8_reduce(mn, OpMin());
  mx = group8_reduce(mx, OpMax());
  const OqQParams q = oq_qparams(mn, mx, 1.f, 1.f, 0, a.n_bits, a.symmetric, a.recip_div);
  if (writer) {
    if (a.scale_out) a.scale_out[r] = q.scale;
    if (a.zp_out) a.zp_out[r] = q.zp;
  }
  const float inv = OQ_DIV(1.0f, q.scale);
  const int inv_ok = oq_inv_ok(inv);
  aq_quant_elems<E>(v, q, inv, inv_ok);
}

template <int E>
__device__ __forceinline__ void aq_row_bwd(const float (&v)[E], float (&g)[E], const AqArgs& a) {
  float mn = INFINITY, mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < E; ++j) { mn = fminf(mn, v[j]); mx = fmaxf(mx, v[j]); }
  mn = group8_reduce(mn, OpMin());
  mx = group8_reduce(mx, OpMax());
  const OqQParams q = oq_qparams(mn, mx, 1.f, 1.f, 0, a.n_bits, a.symmetric, a.recip_div);
  float R = 0.f, nmax = 0.f, nmin = 0.f;
  const float inv = OQ_DIV(1.0f, q.scale);
  const int inv_ok = oq_inv_ok(inv);
  const bool ext = aq_bwd_elems<E>(v, g, q, inv, inv_ok, aq_range(q), mn, mx, R, nmax, nmin);
  R = group8_reduce(R, OpSum());
  nmax = group8_reduce(nmax, OpSum());
  nmin = group8_reduce(nmin, OpSum());
  float gmaxp, gminp;
  oq_scale_grad(R, mn, mx, 1.f, 1.f, 0, a.n_bits, a.symmetric, a.recip_div, &gmaxp, &gminp);
  const float cmax = gmaxp / nmax, cmin = gminp / nmin;
  if (ext) {
#pragma unroll
    for (int j = 0; j < E; ++j) {
      if (v[j] == mx) g[j] += cmax;
      if (v[j] == mn) g[j] += cmin;
    }
  }
}

template <int HALF, bool BWD>
__global__ void __launch_bounds__(256) aq_warp(AqArgs a) {
  constexpr int E = 2 * HALF;
  const int lane = threadIdx.x & 31, sub = lane & 7, slot = lane >> 3;
  const unsigned inner = (unsigned)a.xi.rows_inner;
  const unsigned chunks = (inner + AQ_HPW - 1) / AQ_HPW;
  const unsigned tokens = (unsigned)(a.xi.rows / inner);
  const unsigned total = tokens * chunks, nw = gridDim.x * (blockDim.x >> 5);
  const int c0 = sub * HALF, c1 = a.C / 2 + sub * HALF;       // the lane's two column ranges
  for (unsigned w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < total; w += nw) {
    const unsigned tok = w / chunks, h0 = (w - tok * chunks) * AQ_HPW;
    const unsigned h1 = min(h0 + AQ_HPW, inner);
    float cs[E], sn[E];
    if (a.cos && (int)h0 < a.rope_inner) {
      const unsigned pos = tok % (unsigned)a.seq_len;
      const float* cp = a.cos + (size_t)pos * a.C;
      const float* sp = a.sin + (size_t)pos * a.C;
#pragma unroll
      for (int j = 0; j < HALF; j += 4) {
        *reinterpret_cast<float4*>(cs + j) = *reinterpret_cast<const float4*>(cp + c0 + j);
        *reinterpret_cast<float4*>(cs + HALF + j) = *reinterpret_cast<const float4*>(cp + c1 + j);
        *reinterpret_cast<float4*>(sn + j) = *reinterpret_cast<const float4*>(sp + c0 + j);
        *reinterpret_cast<float4*>(sn + HALF + j) = *reinterpret_cast<const float4*>(sp + c1 + j);
      }
    }
    const bf16* xp = a.x + (size_t)tok * a.xi.ld_outer;
    const bf16* gp = BWD ? a.gy + (size_t)tok * a.gi.ld_outer : nullptr;
    bf16* op = BWD ? a.gx + (size_t)tok * a.gxi.ld_outer : a.y + (size_t)tok * a.yi.ld_outer;
    const long ldo = BWD ? a.gxi.ld_inner : a.yi.ld_inner;
#pragma unroll 2
    for (unsigned hb = h0; hb < h1; hb += 4) {
      const unsigned h = hb + slot;
      const bool live = h < h1;
      const unsigned hh = live ? h : h1 - 1;      // idle row slots shadow the last row (loads stay in bounds, no store)
      float v[E], g[BWD ? E : 1];
      load_half<HALF>(xp + (size_t)hh * a.xi.ld_inner + c0, v);
      load_half<HALF>(xp + (size_t)hh * a.xi.ld_inner + c1, v + HALF);
      if constexpr (BWD) {
        load_half<HALF>(gp + (size_t)hh * a.gi.ld_inner + c0, g);
        load_half<HALF>(gp + (size_t)hh * a.gi.ld_inner + c1, g + HALF);
      }
      const bool rot = a.cos && (int)hh < a.rope_inner;
      if (rot) rope_fwd<HALF>(v, cs, sn);
      if constexpr (BWD) {
        if (a.n_bits < 16) aq_row_bwd<E>(v, g, a);
        if (rot) rope_bwd<HALF>(g, cs, sn);
        if (live) {
          store_half<HALF>(op + (size_t)h * ldo + c0, g);
          store_half<HALF>(op + (size_t)h * ldo + c1, g + HALF);
        }
      } else {
        if (a.n_bits < 16) aq_row_fwd<E>(v, a, (long)tok * inner + hh, live && sub == 0);
        if (live) {
          store_half<HALF>(op + (size_t)h * ldo + c0, v);
          store_half<HALF>(op + (size_t)h * ldo + c1, v + HALF);
        }
      }
    }
  }
}

bool aligned(const void* p, int b) { return ((uintptr_t)p % b) == 0; }

int check_ix(const char* fn, long rows, int rows_inner, int C) {
  OQ_REQUIRE(rows > 0 && C > 0 && rows_inner >= 1 && rows % rows_inner == 0, "%s: bad rows=%ld rows_inner=%d C=%d", fn,
             rows, rows_inner, C);
  return 0;
}

}  // namespace

extern "C" int oq_act_quant_fwd(const void* x, long rows, int C, int rows_inner, long ld_outer, long ld_inner,
                                int n_bits, int symmetric, int recip_div, const float* rope_cos,
                                const float* rope_sin, int seq_len, int rope_inner, void* y, long ldy_outer,
                                long ldy_inner, float* scale_out, float* zp_out, int8_t* codes_s8, long ld_s8,
                                float* s8_meta, cudaStream_t stream) {
  if (check_ix("oq_act_quant_fwd", rows, rows_inner, C)) return -1;
  OQ_REQUIRE((codes_s8 == nullptr) == (s8_meta == nullptr), "oq_act_quant_fwd: codes_s8 and s8_meta go together");
  OQ_REQUIRE(!codes_s8 || (rows_inner == 1 && !rope_cos && n_bits <= 8 && C % 8 == 0 && C <= 12288 && ld_s8 % 8 == 0 &&
                           aligned(codes_s8, 8) && aligned(x, 16) && aligned(y, 16) && ld_outer % 8 == 0 && ldy_outer % 8 == 0),
             "oq_act_quant_fwd: the s8 side output needs plain 2-D rows, C %% 8 == 0, C <= 12288, 8-byte aligned code rows");
  OQ_REQUIRE(x && y, "oq_act_quant_fwd: null pointer");
  OQ_REQUIRE((n_bits >= 2 && n_bits <= 8) || n_bits >= 16, "oq_act_quant_fwd: n_bits %d unsupported", n_bits);
  OQ_REQUIRE(n_bits < 16 || rope_cos, "oq_act_quant_fwd: nothing to do (n_bits>=16 and no rope)");
  OQ_REQUIRE((rope_cos == nullptr) == (rope_sin == nullptr), "oq_act_quant_fwd: need both rope tables");
  OQ_REQUIRE(aligned(rope_cos, 16) && aligned(rope_sin, 16), "oq_act_quant_fwd: rope tables must be 16-byte aligned");
  AqArgs a = {};
  a.x = (const bf16*)x; a.xi = {rows, rows_inner, ld_outer, ld_inner};
  a.y = (bf16*)y; a.yi = {rows, rows_inner, ldy_outer, ldy_inner};
  a.C = C; a.n_bits = n_bits; a.symmetric = symmetric; a.recip_div = recip_div;
  a.cos = rope_cos; a.sin = rope_sin; a.seq_len = seq_len > 0 ? seq_len : 1;
  a.rope_inner = rope_inner > 0 ? rope_inner : rows_inner;
  a.scale_out = scale_out; a.zp_out = zp_out;
  a.s8 = {codes_s8, ld_s8, s8_meta, rows};
  const bool al8 = aligned(x, 16) && aligned(y, 16) && ld_outer % 8 == 0 && ld_inner % 8 == 0 && ldy_outer % 8 == 0 &&
                   ldy_inner % 8 == 0;
  if ((C == 128 || C == 64) && al8 && !codes_s8) {
    OQ_REQUIRE(rows < (1L << 31), "oq_act_quant_fwd: too many rows");
    const long warps = (rows / rows_inner) * ((rows_inner + AQ_HPW - 1) / AQ_HPW);
    const int grid = (int)std::min<long>((warps + 7) / 8, (long)oq_num_sms() * 32);
    if (C == 128) aq_warp<8, false><<<grid, 256, 0, stream>>>(a);
    else aq_warp<4, false><<<grid, 256, 0, stream>>>(a);
  } else {
    OQ_REQUIRE(!rope_cos, "oq_act_quant_fwd: fused rope needs head_dim 64 or 128 and 16-byte aligned rows");
    OQ_REQUIRE(rows < (1L << 31), "oq_act_quant_fwd: too many rows");
    OQ_REQUIRE(!codes_s8 || al8, "oq_act_quant_fwd: the s8 side output needs 16-byte aligned rows");
    const unsigned g = (unsigned)rows;
    if (C % 8 == 0 && al8 && C <= 2048) aq_fwd_block_reg<1><<<g, 256, 0, stream>>>(a);
    else if (C % 8 == 0 && al8 && C <= 4096) aq_fwd_block_reg<2><<<g, 256, 0, stream>>>(a);
    else if (C % 8 == 0 && al8 && C <= 8192) aq_fwd_block_reg<4><<<g, 256, 0, stream>>>(a);
    else if (C % 8 == 0 && al8 && C <= 12288) aq_fwd_block_reg<6><<<g, 256, 0, stream>>>(a);
    else if (C % 8 == 0 && al8) aq_fwd_block<true><<<g, 256, 0, stream>>>(a);
    else aq_fwd_block<false><<<g, 256, 0, stream>>>(a);
  }
  return oq_check_launch("oq_act_quant_fwd");
}

extern "C" int oq_act_quant_bwd(const void* x, long rows, int C, int rows_inner, long ld_outer, long ld_inner,
                                int n_bits, int symmetric, int recip_div, const float* rope_cos,
                                const float* rope_sin, int seq_len, int rope_inner, const void* gy, long ldg_outer,
                                long ldg_inner, void* gx, long ldgx_outer, long ldgx_inner, cudaStream_t stream) {
  if (check_ix("oq_act_quant_bwd", rows, rows_inner, C)) return -1;
  OQ_REQUIRE(x && gy && gx, "oq_act_quant_bwd: null pointer");
  OQ_REQUIRE((n_bits >= 2 && n_bits <= 8) || n_bits >= 16, "oq_act_quant_bwd: n_bits %d unsupported", n_bits);
  OQ_REQUIRE(n_bits < 16 || rope_cos, "oq_act_quant_bwd: nothing to do (n_bits>=16 and no rope)");
  AqArgs a = {};
  a.x = (const bf16*)x; a.xi = {rows, rows_inner, ld_outer, ld_inner};
  OQ_REQUIRE(aligned(rope_cos, 16) && aligned(rope_sin, 16), "oq_act_quant_bwd: rope tables must be 16-byte aligned");
  a.gy = (const bf16*)gy; a.gi = {rows, rows_inner, ldg_outer, ldg_inner};
  a.gx = (bf16*)gx; a.gxi = {rows, rows_inner, ldgx_outer, ldgx_inner};
  a.C = C; a.n_bits = n_bits; a.symmetric = symmetric; a.recip_div = recip_div;
  a.cos = rope_cos; a.sin = rope_sin; a.seq_len = seq_len > 0 ? seq_len : 1;
  a.rope_inner = rope_inner > 0 ? rope_inner : rows_inner;
  const bool al8 = aligned(x, 16) && aligned(gy, 16) && aligned(gx, 16) && ld_outer % 8 == 0 && ld_inner % 8 == 0 &&
                   ldg_outer % 8 == 0 && ldg_inner % 8 == 0 && ldgx_outer % 8 == 0 && ldgx_inner % 8 == 0;
  if ((C == 128 || C == 64) && al8) {
    OQ_REQUIRE(rows < (1L << 31), "oq_act_quant_bwd: too many rows");
    const long warps = (rows / rows_inner) * ((rows_inner + AQ_HPW - 1) / AQ_HPW);
    const int grid = (int)std::min<long>((warps + 7) / 8, (long)oq_num_sms() * 32);
    if (C == 128) aq_warp<8, true><<<grid, 256, 0, stream>>>(a);
    else aq_warp<4, true><<<grid, 256, 0, stream>>>(a);
  } else {
    OQ_REQUIRE(!rope_cos, "oq_act_quant_bwd: fused rope needs head_dim 64 or 128 and 16-byte aligned rows");
    OQ_REQUIRE(rows < (1L << 31), "oq_act_quant_bwd: too many rows");
    const unsigned g = (unsigned)rows;
    if (C % 8 == 0 && al8 && C <= 2048) aq_bwd_block_reg<1><<<g, 256, 0, stream>>>(a);
    else if (C % 8 == 0 && al8 && C <= 4096) aq_bwd_block_reg<2><<<g, 256, 0, stream>>>(a);
    else if (C % 8 == 0 && al8 && C <= 8192) aq_bwd_block_reg<4><<<g, 256, 0, stream>>>(a);
    else if (C % 8 == 0 && al8 && C <= 12288) aq_bwd_block_reg<6><<<g, 256, 0, stream>>>(a);
    else if (C % 8 == 0 && al8) aq_bwd_block<true><<<g, 256, 0, stream>>>(a);
    else aq_bwd_block<false><<<g, 256, 0, stream>>>(a);
  }
  return oq_check_launch("oq_act_quant_bwd");
}

/* SwiGLU + per-token fake-quant of the result (LlamaMLP act_fn(gate)*up feeding down_proj's act_quantizer,
 * models/int_llama_layer.py:44-45 + quantize/int_linear.py:59-60). */
extern "C" int oq_swiglu_act_quant_fwd(const void* gu, long ldgu, long rows, int F, int n_bits, int symmetric,
                                       int recip_div, void* mq, long ldmq, int8_t* codes_s8, long ld_s8, float* s8_meta,
                                       cudaStream_t stream) {
  OQ_REQUIRE((codes_s8 == nullptr) == (s8_meta == nullptr), "oq_swiglu_act_quant_fwd: codes_s8 and s8_meta go together");
  OQ_REQUIRE(!codes_s8 || (ld_s8 % 8 == 0 && aligned(codes_s8, 8)), "oq_swiglu_act_quant_fwd: code rows must be 8-byte aligned");
  OQ_REQUIRE(gu && mq && rows > 0 && rows < (1L << 31) && F > 0, "oq_swiglu_act_quant_fwd: null pointer or bad shape");
  OQ_REQUIRE(n_bits >= 2 && n_bits <= 8, "oq_swiglu_act_quant_fwd: n_bits %d unsupported", n_bits);
  OQ_REQUIRE(F % 8 == 0 && F <= 16384 && ldgu % 8 == 0 && ldmq % 8 == 0 && aligned(gu, 16) && aligned(mq, 16),
             "oq_swiglu_act_quant_fwd: need F %% 8 == 0, F <= 16384, strides multiples of 8, 16-byte aligned pointers");
  GluArgs a = {};
  a.gu = (const bf16*)gu; a.ldgu = ldgu; a.mq = (bf16*)mq; a.ldmq = ldmq; a.F = F; a.n_bits = n_bits;
  a.symmetric = symmetric; a.recip_div = recip_div;
  a.s8 = {codes_s8, ld_s8, s8_meta, rows};
  const unsigned g = (unsigned)rows;
  if (F <= 4096) swiglu_aq_fwd_kernel<1><<<g, 512, 0, stream>>>(a);
  else if (F <= 8192) swiglu_aq_fwd_kernel<2><<<g, 512, 0, stream>>>(a);
  else if (F <= 12288) swiglu_aq_fwd_kernel<3><<<g, 512, 0, stream>>>(a);
  else swiglu_aq_fwd_kernel<4><<<g, 512, 0, stream>>>(a);
  return oq_check_launch("oq_swiglu_act_quant_fwd");
}

extern "C" int oq_swiglu_act_quant_bwd(const void* gu, long ldgu, long rows, int F, int n_bits, int symmetric,
                                       int recip_div, const void* gy, long ldgy, void* dgu, long lddgu,
                                       cudaStream_t stream) {
  OQ_REQUIRE(gu && gy && dgu && rows > 0 && rows < (1L << 31) && F > 0, "oq_swiglu_act_quant_bwd: null pointer or bad shape");
  OQ_REQUIRE(n_bits >= 2 && n_bits <= 8, "oq_swiglu_act_quant_bwd: n_bits %d unsupported", n_bits);
  OQ_REQUIRE(F % 8 == 0 && F <= 16384 && ldgu % 8 == 0 && ldgy % 8 == 0 && lddgu % 8 == 0 && aligned(gu, 16) &&
                 aligned(gy, 16) && aligned(dgu, 16),
             "oq_swiglu_act_quant_bwd: need F %% 8 == 0, F <= 16384, strides multiples of 8, 16-byte aligned pointers");
  GluArgs a = {};
  a.gu = (const bf16*)gu; a.ldgu = ldgu; a.gy = (const bf16*)gy; a.ldgy = ldgy; a.dgu = (bf16*)dgu; a.lddgu = lddgu;
  a.F = F; a.n_bits = n_bits; a.symmetric = symmetric; a.recip_div = recip_div;
  const unsigned g = (unsigned)rows;
  static bool attr = false;
  if (!attr) {      // the parked rows of the two largest variants exceed the 48 KB default
    if (cudaFuncSetAttribute(swiglu_aq_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 3 * 512 * 16) != cudaSuccess ||
        cudaFuncSetAttribute(swiglu_aq_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4 * 512 * 16) != cudaSuccess)
      return oq_check_launch("oq_swiglu_act_quant_bwd (shared memory attribute)");
    attr = true;
  }
  if (F <= 4096) swiglu_aq_bwd_kernel<1><<<g, 512, 2 * 1 * 512 * 16, stream>>>(a);
  else if (F <= 8192) swiglu_aq_bwd_kernel<2><<<g, 512, 2 * 2 * 512 * 16, stream>>>(a);
  else if (F <= 12288) swiglu_aq_bwd_kernel<3><<<g, 512, 2 * 3 * 512 * 16, stream>>>(a);
  else swiglu_aq_bwd_kernel<4><<<g, 512, 2 * 4 * 512 * 16, stream>>>(a);
  return oq_check_launch("oq_swiglu_act_quant_bwd");
}
