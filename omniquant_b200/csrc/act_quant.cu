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
  __device__ __forceinline__ long off(long r) const { return (r / rows_inner) * ld_outer + (r % rows_inner) * ld_inner; }
};

struct AqArgs {
  const bf16* x; RowIx xi;
  bf16* y; RowIx yi;
  const bf16* gy; RowIx gi;
  bf16* gx; RowIx gxi;
  int C, n_bits, symmetric, recip_div;
  const float* cos; const float* sin; int seq_len;  // rope tables [seq_len, C] or null
  float* scale_out; float* zp_out;
};

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
  __shared__ float scratch[2 * 8];
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
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (threadIdx.x + i * 256) * 8;
    if (c < a.C) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = oq_clampf(OQ_ADD(oq_rint_div(v[i][j], q.scale, inv, inv_ok), q.zp), 0.0f, q.qmax);
        v[i][j] = oq_dequant(t, q);
      }
      store_bf8(yr + c, v[i]);
    }
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
  const float inv = OQ_DIV(1.0f, q.scale);
  const int inv_ok = oq_inv_ok(inv);
  float R = 0.f, nmax = 0.f, nmin = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (threadIdx.x + i * 256) * 8;
    if (c < a.C) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float x = v[i][j];
        const float qa = x * inv;
        const float t = OQ_ADD(oq_rint_div(x, q.scale, inv, inv_ok), q.zp);   // same codes as the forward, ties included
        const bool inr = (t >= 0.0f) && (t <= q.qmax);
        const float cc = oq_clampf(t, 0.0f, q.qmax);
        R += g[i][j] * ((cc - q.zp) - (inr ? qa : 0.0f));
        nmax += (x == mx) ? 1.0f : 0.0f;
        nmin += (x == mn) ? 1.0f : 0.0f;
        if (!inr) g[i][j] = 0.0f;          // straight-through part
      }
    }
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
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (v[i][j] == mx) g[i][j] += cmax;
        if (v[i][j] == mn) g[i][j] += cmin;
      }
      store_bf8(dr + c, g[i]);
    }
  }
}

// ---- warp per row (C = 32*EPL, optional rope) --------------------------------
template <int EPL>
__device__ __forceinline__ void load_row(const bf16* p, float (&v)[EPL]) {
  if constexpr (EPL == 8) load_bf8(p, v);
  else if constexpr (EPL == 4) load_bf4(p, v);
  else if constexpr (EPL == 2) { uint32_t u = *reinterpret_cast<const uint32_t*>(p); unpack_bf2(u, v[0], v[1]); }
  else {
#pragma unroll
    for (int j = 0; j < EPL; ++j) v[j] = bf2f(p[j]);
  }
}
template <int EPL>
__device__ __forceinline__ void store_row(bf16* p, const float (&v)[EPL]) {
  if constexpr (EPL == 8) store_bf8(p, v);
  else if constexpr (EPL == 4) store_bf4(p, v);
  else if constexpr (EPL == 2) *reinterpret_cast<uint32_t*>(p) = pack_bf2(v[0], v[1]);
  else {
#pragma unroll
    for (int j = 0; j < EPL; ++j) p[j] = f2bf(v[j]);
  }
}

// y = x*cos + rotate_half(x)*sin ; rotate_half(x)[i] = -x[i+h] (i<h), x[i-h] (i>=h)
// (transformers 4.31 apply_rotary_pos_emb as used at models/int_llama_layer.py:125)
template <int EPL>
__device__ __forceinline__ void rope_fwd(float (&v)[EPL], const float* cs, const float* sn, int lane) {
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    const float p = __shfl_xor_sync(0xffffffffu, v[j], 16);
    const float rot = lane < 16 ? -p : p;
    v[j] = v[j] * cs[j] + rot * sn[j];
  }
}
// transpose of the above: dx[j] = g[j]cos[j] + (j<h ? g[j+h]sin[j+h] : -g[j-h]sin[j-h])
template <int EPL>
__device__ __forceinline__ void rope_bwd(float (&g)[EPL], const float* cs, const float* sn, int lane) {
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    const float gs = g[j] * sn[j];
    const float p = __shfl_xor_sync(0xffffffffu, gs, 16);
    g[j] = g[j] * cs[j] + (lane < 16 ? p : -p);
  }
}

template <int EPL>
__global__ void __launch_bounds__(256) aq_fwd_warp(AqArgs a) {
  const int lane = threadIdx.x & 31;
  const long wpb = blockDim.x >> 5;
  constexpr int RB = 1;   // rows in flight per warp (measured: 4 is slower, the kernel is ALU-bound on the IEEE divisions)
  const long nw = (long)gridDim.x * wpb;
  for (long r0 = blockIdx.x * wpb + (threadIdx.x >> 5); r0 < a.xi.rows; r0 += nw * RB) {
    float vv[RB][EPL];
#pragma unroll
    for (int u = 0; u < RB; ++u)
      if (r0 + u * nw < a.xi.rows) load_row<EPL>(a.x + a.xi.off(r0 + u * nw) + lane * EPL, vv[u]);
#pragma unroll
    for (int u = 0; u < RB; ++u) {
    const long r = r0 + u * nw;
    if (r >= a.xi.rows) break;
    float (&v)[EPL] = vv[u];
    if (a.cos) {
      const long pos = (r / a.xi.rows_inner) % a.seq_len;
      float cs[EPL], sn[EPL];
#pragma unroll
      for (int j = 0; j < EPL; ++j) {
        cs[j] = a.cos[pos * a.C + lane * EPL + j];
        sn[j] = a.sin[pos * a.C + lane * EPL + j];
      }
      rope_fwd<EPL>(v, cs, sn, lane);
    }
    if (a.n_bits < 16) {
      float mn = INFINITY, mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < EPL; ++j) { mn = fminf(mn, v[j]); mx = fmaxf(mx, v[j]); }
      mn = warp_reduce(mn, OpMin());
      mx = warp_reduce(mx, OpMax());
      const OqQParams q = oq_qparams(mn, mx, 1.f, 1.f, 0, a.n_bits, a.symmetric, a.recip_div);
      if (lane == 0) {
        if (a.scale_out) a.scale_out[r] = q.scale;
        if (a.zp_out) a.zp_out[r] = q.zp;
      }
      const float inv = OQ_DIV(1.0f, q.scale);
      const int inv_ok = oq_inv_ok(inv);
#pragma unroll
      for (int j = 0; j < EPL; ++j)
        v[j] = oq_dequant(oq_clampf(OQ_ADD(oq_rint_div(v[j], q.scale, inv, inv_ok), q.zp), 0.0f, q.qmax), q);
    }
    store_row<EPL>(a.y + a.yi.off(r) + lane * EPL, v);
    }
  }
}

template <int EPL>
__global__ void __launch_bounds__(256) aq_bwd_warp(AqArgs a) {
  const int lane = threadIdx.x & 31;
  const long wpb = blockDim.x >> 5;
  constexpr int RB = 1;
  const long nw = (long)gridDim.x * wpb;
  for (long r0 = blockIdx.x * wpb + (threadIdx.x >> 5); r0 < a.xi.rows; r0 += nw * RB) {
    float vv[RB][EPL], gg[RB][EPL];
#pragma unroll
    for (int u = 0; u < RB; ++u)
      if (r0 + u * nw < a.xi.rows) {
        load_row<EPL>(a.x + a.xi.off(r0 + u * nw) + lane * EPL, vv[u]);
        load_row<EPL>(a.gy + a.gi.off(r0 + u * nw) + lane * EPL, gg[u]);
      }
#pragma unroll
    for (int u = 0; u < RB; ++u) {
    const long r = r0 + u * nw;
    if (r >= a.xi.rows) break;
    float (&v)[EPL] = vv[u];
    float (&g)[EPL] = gg[u];
    float cs[EPL], sn[EPL];
    if (a.cos) {
      const long pos = (r / a.xi.rows_inner) % a.seq_len;
#pragma unroll
      for (int j = 0; j < EPL; ++j) {
        cs[j] = a.cos[pos * a.C + lane * EPL + j];
        sn[j] = a.sin[pos * a.C + lane * EPL + j];
      }
      rope_fwd<EPL>(v, cs, sn, lane);
    }
    if (a.n_bits < 16) {
      float mn = INFINITY, mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < EPL; ++j) { mn = fminf(mn, v[j]); mx = fmaxf(mx, v[j]); }
      mn = warp_reduce(mn, OpMin());
      mx = warp_reduce(mx, OpMax());
      const OqQParams q = oq_qparams(mn, mx, 1.f, 1.f, 0, a.n_bits, a.symmetric, a.recip_div);
      float R = 0.f, nmax = 0.f, nmin = 0.f;
      bool inr[EPL];
      const float inv = OQ_DIV(1.0f, q.scale);
      const int inv_ok = oq_inv_ok(inv);
#pragma unroll
      for (int j = 0; j < EPL; ++j) {
        const float qa = v[j] * inv;
        const float t = OQ_ADD(oq_rint_div(v[j], q.scale, inv, inv_ok), q.zp);   // same codes as the forward, ties included
        inr[j] = (t >= 0.0f) && (t <= q.qmax);
        const float cc = oq_clampf(t, 0.0f, q.qmax);
        R += g[j] * ((cc - q.zp) - (inr[j] ? qa : 0.0f));
        nmax += (v[j] == mx) ? 1.0f : 0.0f;
        nmin += (v[j] == mn) ? 1.0f : 0.0f;
      }
      R = warp_reduce(R, OpSum());
      nmax = warp_reduce(nmax, OpSum());
      nmin = warp_reduce(nmin, OpSum());
      float gmaxp, gminp;
      oq_scale_grad(R, mn, mx, 1.f, 1.f, 0, a.n_bits, a.symmetric, a.recip_div, &gmaxp, &gminp);
      const float cmax = gmaxp / nmax, cmin = gminp / nmin;
#pragma unroll
      for (int j = 0; j < EPL; ++j) {
        float d = inr[j] ? g[j] : 0.0f;
        if (v[j] == mx) d += cmax;
        if (v[j] == mn) d += cmin;
        g[j] = d;
      }
    }
    if (a.cos) rope_bwd<EPL>(g, cs, sn, lane);
    store_row<EPL>(a.gx + a.gxi.off(r) + lane * EPL, g);
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
                                const float* rope_sin, int seq_len, void* y, long ldy_outer, long ldy_inner,
                                float* scale_out, float* zp_out, cudaStream_t stream) {
  if (check_ix("oq_act_quant_fwd", rows, rows_inner, C)) return -1;
  OQ_REQUIRE(x && y, "oq_act_quant_fwd: null pointer");
  OQ_REQUIRE((n_bits >= 2 && n_bits <= 8) || n_bits >= 16, "oq_act_quant_fwd: n_bits %d unsupported", n_bits);
  OQ_REQUIRE(n_bits < 16 || rope_cos, "oq_act_quant_fwd: nothing to do (n_bits>=16 and no rope)");
  OQ_REQUIRE((rope_cos == nullptr) == (rope_sin == nullptr), "oq_act_quant_fwd: need both rope tables");
  AqArgs a = {};
  a.x = (const bf16*)x; a.xi = {rows, rows_inner, ld_outer, ld_inner};
  a.y = (bf16*)y; a.yi = {rows, rows_inner, ldy_outer, ldy_inner};
  a.C = C; a.n_bits = n_bits; a.symmetric = symmetric; a.recip_div = recip_div;
  a.cos = rope_cos; a.sin = rope_sin; a.seq_len = seq_len > 0 ? seq_len : 1;
  a.scale_out = scale_out; a.zp_out = zp_out;
  const bool al8 = aligned(x, 16) && aligned(y, 16) && ld_outer % 8 == 0 && ld_inner % 8 == 0 && ldy_outer % 8 == 0 &&
                   ldy_inner % 8 == 0;
  if ((C == 128 || C == 64) && al8) {
    const int grid = (int)std::min<long>((rows + 7) / 8, (long)oq_num_sms() * 8);
    if (C == 128) aq_fwd_warp<4><<<grid, 256, 0, stream>>>(a);
    else aq_fwd_warp<2><<<grid, 256, 0, stream>>>(a);
  } else {
    OQ_REQUIRE(!rope_cos, "oq_act_quant_fwd: fused rope needs head_dim 64 or 128 and 16-byte aligned rows");
    OQ_REQUIRE(rows < (1L << 31), "oq_act_quant_fwd: too many rows");
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
                                const float* rope_sin, int seq_len, const void* gy, long ldg_outer, long ldg_inner,
                                void* gx, long ldgx_outer, long ldgx_inner, cudaStream_t stream) {
  if (check_ix("oq_act_quant_bwd", rows, rows_inner, C)) return -1;
  OQ_REQUIRE(x && gy && gx, "oq_act_quant_bwd: null pointer");
  OQ_REQUIRE((n_bits >= 2 && n_bits <= 8) || n_bits >= 16, "oq_act_quant_bwd: n_bits %d unsupported", n_bits);
  OQ_REQUIRE(n_bits < 16 || rope_cos, "oq_act_quant_bwd: nothing to do (n_bits>=16 and no rope)");
  AqArgs a = {};
  a.x = (const bf16*)x; a.xi = {rows, rows_inner, ld_outer, ld_inner};
  a.gy = (const bf16*)gy; a.gi = {rows, rows_inner, ldg_outer, ldg_inner};
  a.gx = (bf16*)gx; a.gxi = {rows, rows_inner, ldgx_outer, ldgx_inner};
  a.C = C; a.n_bits = n_bits; a.symmetric = symmetric; a.recip_div = recip_div;
  a.cos = rope_cos; a.sin = rope_sin; a.seq_len = seq_len > 0 ? seq_len : 1;
  const bool al8 = aligned(x, 16) && aligned(gy, 16) && aligned(gx, 16) && ld_outer % 8 == 0 && ld_inner % 8 == 0 &&
                   ldg_outer % 8 == 0 && ldg_inner % 8 == 0 && ldgx_outer % 8 == 0 && ldgx_inner % 8 == 0;
  if ((C == 128 || C == 64) && al8) {
    const int grid = (int)std::min<long>((rows + 7) / 8, (long)oq_num_sms() * 8);
    if (C == 128) aq_bwd_warp<4><<<grid, 256, 0, stream>>>(a);
    else aq_bwd_warp<2><<<grid, 256, 0, stream>>>(a);
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
