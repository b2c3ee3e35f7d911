// K7: OmniLlamaRMSNorm / OmniLayerNorm forward + backward with the step's
// temporary (LET-folded) weight and bias.
//
// Reference: quantize/omni_norm.py:52-63 (RMS: fp32 variance, (w*x_hat + b).to(dtype))
// and :26-33 (F.layer_norm).  The backward also produces the per-channel sums
// dL/dw', dL/db' that the LET scale/shift gradients are built from
// (w' = w/s, b' = (b - delta)/s; models/transformation.py:28-33).
//
// HBM-bound: forward 2 B read + 2 B write per element; backward 2+2 B read,
// 2 B write.  One 256-thread block per token row, fp32 math, bf16 storage.
#include <algorithm>

#include "common.cuh"

namespace {

struct NormArgs {
  const bf16* x; long ldx;
  long rows; int C;
  const float* w; const float* b;
  float eps; int kind;  // 0 rms, 1 layernorm
  bf16* y; long ldy;
  float* mean; float* rstd;
  const bf16* dy; long lddy;
  bf16* dx; long lddx;
  const bf16* dx_add; long ldadd;
  float* dw_acc; float* db_acc;
  bf16* yq; long ldyq;      // optional fused per-token fake-quant of y (n_bits < 16)
  int n_bits, symmetric, recip_div;
  S8Side s8;                // optional exact-integer side output of the fused quantiser
};

template <bool VEC>
__global__ void __launch_bounds__(256) norm_fwd_kernel(NormArgs a) {
  __shared__ float scratch[2 * 32];
  constexpr int V = VEC ? 8 : 1;
  const long r = blockIdx.x;
  const bf16* xr = a.x + r * a.ldx;
  bf16* yr = a.y + r * a.ldy;
  float s[2] = {0.f, 0.f};
  for (int c = threadIdx.x * V; c < a.C; c += blockDim.x * V) {
    float v[V];
    if constexpr (VEC) load_bf8(xr + c, v);
    else v[0] = bf2f(xr[c]);
#pragma unroll
    for (int j = 0; j < V; ++j) { s[0] += v[j]; s[1] += v[j] * v[j]; }
  }
  block_reduce<2>(s, scratch, OpSum(), 0.0f);
  float mean = 0.f, rstd;
  if (a.kind == 1) {
    mean = s[0] / a.C;
    // two-pass variance for accuracy (matches native_layer_norm's Welford result closely)
    float vs[1] = {0.f};
    for (int c = threadIdx.x * V; c < a.C; c += blockDim.x * V) {
      float v[V];
      if constexpr (VEC) load_bf8(xr + c, v);
      else v[0] = bf2f(xr[c]);
#pragma unroll
      for (int j = 0; j < V; ++j) { const float d = v[j] - mean; vs[0] += d * d; }
    }
    block_reduce<1>(vs, scratch, OpSum(), 0.0f);
    rstd = rsqrtf(vs[0] / a.C + a.eps);
  } else {
    rstd = rsqrtf(s[1] / a.C + a.eps);
  }
  if (threadIdx.x == 0) {
    if (a.mean) a.mean[r] = mean;
    if (a.rstd) a.rstd[r] = rstd;
  }
  for (int c = threadIdx.x * V; c < a.C; c += blockDim.x * V) {
    float v[V];
    if constexpr (VEC) load_bf8(xr + c, v);
    else v[0] = bf2f(xr[c]);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float h = (v[j] - mean) * rstd;
      v[j] = a.w[c + j] * h + (a.b ? a.b[c + j] : 0.0f);
    }
    if constexpr (VEC) store_bf8(yr + c, v);
    else yr[c] = f2bf(v[0]);
  }
}

// Register-resident forward (C <= TPB*8*NV, C % 8 == 0): one read of x, one write of y; optionally also writes
// yq = per-token fake-quant of the bf16-rounded y (QuantLinear.act_quantizer on the norm output,
// quantize/int_linear.py:59-60) -- bit-identical to oq_act_quant_fwd run on y afterwards.
template <int NV, int TPB>
__global__ void __launch_bounds__(TPB) norm_fwd_reg(NormArgs a) {
  __shared__ float scratch[2 * 32];
  const long r = blockIdx.x;
  const bf16* xr = a.x + r * a.ldx;
  float v[NV][8];
  float s[2] = {0.f, 0.f};
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (threadIdx.x + i * TPB) * 8;
    if (c < a.C) {
      load_bf8(xr + c, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[0] += v[i][j]; s[1] += v[i][j] * v[i][j]; }
    }
  }
  block_reduce<2>(s, scratch, OpSum(), 0.0f);
  float mean = 0.f, rstd;
  if (a.kind == 1) {
    mean = s[0] / a.C;
    float vs[1] = {0.f};
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (threadIdx.x + i * TPB) * 8;
      if (c < a.C) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; vs[0] += d * d; }
      }
    }
    block_reduce<1>(vs, scratch, OpSum(), 0.0f);
    rstd = rsqrtf(vs[0] / a.C + a.eps);
  } else {
    rstd = rsqrtf(s[1] / a.C + a.eps);
  }
  if (threadIdx.x == 0) {
    if (a.mean) a.mean[r] = mean;
    if (a.rstd) a.rstd[r] = rstd;
  }
  float mm[2] = {-INFINITY, -INFINITY};   // -min, max of the bf16-rounded output
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (threadIdx.x + i * TPB) * 8;
    if (c < a.C) {
      const float4 w0 = *reinterpret_cast<const float4*>(a.w + c), w1 = *reinterpret_cast<const float4*>(a.w + c + 4);
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      float b[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (a.b) {
        const float4 b0 = *reinterpret_cast<const float4*>(a.b + c), b1 = *reinterpret_cast<const float4*>(a.b + c + 4);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float h = (v[i][j] - mean) * rstd;
        v[i][j] = w[j] * h + b[j];
      }
      if (a.y) store_bf8(a.y + r * a.ldy + c, v[i]);
      if (a.yq) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[i][j] = bf16_round(v[i][j]);
          mm[0] = fmaxf(mm[0], -v[i][j]);
          mm[1] = fmaxf(mm[1], v[i][j]);
        }
      }
    }
  }
  if (!a.yq) return;
  block_reduce<2>(mm, scratch, OpMax(), -INFINITY);
  const OqQParams q = oq_qparams(-mm[0], mm[1], 1.f, 1.f, 0, a.n_bits, a.symmetric, a.recip_div);
  const float inv = OQ_DIV(1.0f, q.scale);
  const bool fast = oq_inv_ok(inv);
  const float z0 = oq_s8_z0(q.zp, q.qmax);
  int isum = 0;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (threadIdx.x + i * TPB) * 8;
    if (c < a.C) {
      float cd[8];
      if (fast) {   // same arithmetic as the stand-alone quantiser (csrc/act_quant.cu aq_t); branch uniform per row
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float x = v[i][j], qa = __fmul_rn(x, inv);
          const float t = OQ_ADD(rintf(__fmaf_rn(__fmaf_rn(-qa, q.scale, x), inv, qa)), q.zp);
          cd[j] = oq_clampf(t, 0.0f, q.qmax);
          v[i][j] = oq_dequant(cd[j], q);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = OQ_ADD(rintf(OQ_DIV(v[i][j], q.scale)), q.zp);
          cd[j] = oq_clampf(t, 0.0f, q.qmax);
          v[i][j] = oq_dequant(cd[j], q);
        }
      }
      store_bf8(a.yq + r * a.ldyq + c, v[i]);
      if (a.s8.codes) s8_store8(a.s8.codes + r * a.s8.ld + c, cd, z0, isum);
    }
  }
  if (a.s8.codes) {
    float csum[1] = {(float)isum};        // exact: |sum| <= 128 * C < 2^24
    block_reduce<1>(csum, scratch, OpSum(), 0.0f);
    if (threadIdx.x == 0) s8_write_meta(a.s8, r, q, z0, csum[0]);
  }
}

// dynamic smem: [C] dw accumulators then [C] db accumulators (when requested)
template <bool VEC>
__global__ void __launch_bounds__(256) norm_bwd_kernel(NormArgs a) {
  extern __shared__ float dsm[];
  __shared__ float scratch[2 * 32];
  constexpr int V = VEC ? 8 : 1;
  float* dwacc = a.dw_acc ? dsm : nullptr;
  float* dbacc = a.db_acc ? dsm + (a.dw_acc ? a.C : 0) : nullptr;
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    if (dwacc) dwacc[c] = 0.f;
    if (dbacc) dbacc[c] = 0.f;
  }
  __syncthreads();
  for (long r = blockIdx.x; r < a.rows; r += gridDim.x) {
    const bf16* xr = a.x + r * a.ldx;
    const bf16* gr = a.dy + r * a.lddy;
    const float mean = a.kind == 1 ? a.mean[r] : 0.f, rstd = a.rstd[r];
    float s[2] = {0.f, 0.f};  // sum(dxhat*xhat), sum(dxhat)
    for (int c = threadIdx.x * V; c < a.C; c += blockDim.x * V) {
      float v[V], g[V];
      if constexpr (VEC) { load_bf8(xr + c, v); load_bf8(gr + c, g); }
      else { v[0] = bf2f(xr[c]); g[0] = bf2f(gr[c]); }
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float h = (v[j] - mean) * rstd;
        const float dh = g[j] * a.w[c + j];
        s[0] += dh * h;
        s[1] += dh;
        if (dwacc) dwacc[c + j] += g[j] * h;
        if (dbacc) dbacc[c + j] += g[j];
      }
    }
    if (a.dx) {
      block_reduce<2>(s, scratch, OpSum(), 0.0f);
      const float m1 = s[0] / a.C, m2 = a.kind == 1 ? s[1] / a.C : 0.f;
      bf16* dr = a.dx + r * a.lddx;
      const bf16* ar = a.dx_add ? a.dx_add + r * a.ldadd : nullptr;
      for (int c = threadIdx.x * V; c < a.C; c += blockDim.x * V) {
        float v[V], g[V], ad[V];
        if constexpr (VEC) {
          load_bf8(xr + c, v); load_bf8(gr + c, g);
          if (ar) load_bf8(ar + c, ad);
        } else {
          v[0] = bf2f(xr[c]); g[0] = bf2f(gr[c]);
          if (ar) ad[0] = bf2f(ar[c]);
        }
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float h = (v[j] - mean) * rstd;
          const float dh = g[j] * a.w[c + j];
          g[j] = rstd * (dh - m2 - h * m1) + (ar ? ad[j] : 0.f);
        }
        if constexpr (VEC) store_bf8(dr + c, g);
        else dr[c] = f2bf(g[0]);
      }
    }
  }
  __syncthreads();  // column owners in the row loop differ from the flush mapping below
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    if (dwacc) atomicAdd(a.dw_acc + c, dwacc[c]);
    if (dbacc) atomicAdd(a.db_acc + c, dbacc[c]);
  }
}

// Register-resident variant (C <= 2048*NV, C % 8 == 0): a thread owns the same 8*NV columns for every row its block
// visits, so the per-channel sums dL/dw', dL/db' accumulate in registers (the shared-memory accumulators of the kernel
// above serialise on 8-way bank conflicts) and x / dy are read once.  Flushed with 16-byte vector atomics.
template <int NV, bool ACC, int TPB>
__global__ void __launch_bounds__(TPB, (NV == 1 ? 1024 : 512) / TPB) norm_bwd_reg(NormArgs a) {
  __shared__ float scratch[2 * 32];
  float wv[NV][8], dw[ACC ? NV : 1][8], db[ACC ? NV : 1][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (threadIdx.x + i * TPB) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      wv[i][j] = c < a.C ? a.w[c + j] : 0.f;
      if constexpr (ACC) { dw[i][j] = 0.f; db[i][j] = 0.f; }
    }
  }
  for (long r = blockIdx.x; r < a.rows; r += gridDim.x) {
    const bf16* xr = a.x + r * a.ldx;
    const bf16* gr = a.dy + r * a.lddy;
    const float mean = a.kind == 1 ? a.mean[r] : 0.f, rstd = a.rstd[r];
    float h[NV][8], dh[NV][8], ad[NV][8];
    float s[2] = {0.f, 0.f};  // sum(dxhat*xhat), sum(dxhat)
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (threadIdx.x + i * TPB) * 8;
      if (c < a.C) {
        load_bf8(xr + c, h[i]);
        load_bf8(gr + c, dh[i]);
        if (a.dx && a.dx_add) load_bf8(a.dx_add + r * a.ldadd + c, ad[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (threadIdx.x + i * TPB) * 8;
      if (c < a.C) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float hh = (h[i][j] - mean) * rstd, g = dh[i][j];
          const float d = g * wv[i][j];
          s[0] += d * hh;
          s[1] += d;
          if constexpr (ACC) { dw[i][j] += g * hh; db[i][j] += g; }
          h[i][j] = hh;
          dh[i][j] = d;
        }
      }
    }
    if (a.dx) {
      block_reduce<2>(s, scratch, OpSum(), 0.0f);
      const float m1 = s[0] / a.C, m2 = a.kind == 1 ? s[1] / a.C : 0.f;
      bf16* dr = a.dx + r * a.lddx;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (threadIdx.x + i * TPB) * 8;
        if (c < a.C) {
#pragma unroll
          for (int j = 0; j < 8; ++j) dh[i][j] = rstd * (dh[i][j] - m2 - h[i][j] * m1) + (a.dx_add ? ad[i][j] : 0.f);
          store_bf8(dr + c, dh[i]);
        }
      }
    }
  }
  if constexpr (ACC) {
    const bool v4 = ((uintptr_t)a.dw_acc % 16 == 0) && ((uintptr_t)a.db_acc % 16 == 0);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (threadIdx.x + i * TPB) * 8;
      if (c < a.C) {
        if (v4) {
          if (a.dw_acc) {
            atomicAdd(reinterpret_cast<float4*>(a.dw_acc + c), make_float4(dw[i][0], dw[i][1], dw[i][2], dw[i][3]));
            atomicAdd(reinterpret_cast<float4*>(a.dw_acc + c + 4), make_float4(dw[i][4], dw[i][5], dw[i][6], dw[i][7]));
          }
          if (a.db_acc) {
            atomicAdd(reinterpret_cast<float4*>(a.db_acc + c), make_float4(db[i][0], db[i][1], db[i][2], db[i][3]));
            atomicAdd(reinterpret_cast<float4*>(a.db_acc + c + 4), make_float4(db[i][4], db[i][5], db[i][6], db[i][7]));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (a.dw_acc) atomicAdd(a.dw_acc + c + j, dw[i][j]);
            if (a.db_acc) atomicAdd(a.db_acc + c + j, db[i][j]);
          }
        }
      }
    }
  }
}

}  // namespace

namespace {
int norm_fwd_launch(const char* fn, NormArgs a, cudaStream_t stream) {
  const bool al = (uintptr_t)a.x % 16 == 0 && (uintptr_t)a.y % 16 == 0 && (uintptr_t)a.yq % 16 == 0 &&
                  (uintptr_t)a.w % 16 == 0 && (uintptr_t)a.b % 16 == 0;
  const bool vec = a.C % 8 == 0 && a.ldx % 8 == 0 && (!a.y || a.ldy % 8 == 0) && (!a.yq || a.ldyq % 8 == 0) && al;
  const unsigned g = (unsigned)a.rows;
  if (vec && a.C <= 2048) norm_fwd_reg<1, 256><<<g, 256, 0, stream>>>(a);
  else if (vec && a.C <= 4096) norm_fwd_reg<2, 256><<<g, 256, 0, stream>>>(a);   // 16 elements per thread amortise the reductions
  else if (vec && a.C <= 8192) norm_fwd_reg<2, 512><<<g, 512, 0, stream>>>(a);
  else if (vec && a.C <= 16384) norm_fwd_reg<4, 512><<<g, 512, 0, stream>>>(a);
  else {
    if (a.yq) {
      oq_set_error("%s: the fused quantiser needs C %% 8 == 0, C <= 16384, strides multiples of 8, 16-byte aligned pointers", fn);
      return -1;
    }
    if (vec) norm_fwd_kernel<true><<<g, 256, 0, stream>>>(a);
    else norm_fwd_kernel<false><<<g, 256, 0, stream>>>(a);
  }
  return oq_check_launch(fn);
}
}  // namespace

extern "C" int oq_norm_fwd(const void* x, long rows, int C, long ldx, const float* w, const float* b, float eps,
                           int kind, void* y, long ldy, float* mean_out, float* rstd_out, cudaStream_t stream) {
  OQ_REQUIRE(x && y && w && rows > 0 && C > 0, "oq_norm_fwd: null pointer or bad shape");
  OQ_REQUIRE(kind == 0 || kind == 1, "oq_norm_fwd: kind must be 0 (rms) or 1 (layernorm)");
  OQ_REQUIRE(rows < (1L << 31), "oq_norm_fwd: too many rows");
  NormArgs a = {};
  a.x = (const bf16*)x; a.ldx = ldx; a.rows = rows; a.C = C; a.w = w; a.b = b; a.eps = eps; a.kind = kind;
  a.y = (bf16*)y; a.ldy = ldy; a.mean = mean_out; a.rstd = rstd_out;
  return norm_fwd_launch("oq_norm_fwd", a, stream);
}

extern "C" int oq_norm_act_quant_fwd(const void* x, long rows, int C, long ldx, const float* w, const float* b, float eps,
                                     int kind, void* y, long ldy, float* mean_out, float* rstd_out, int n_bits,
                                     int symmetric, int recip_div, void* yq, long ldyq, int8_t* codes_s8, long ld_s8,
                                     float* s8_meta, cudaStream_t stream) {
  OQ_REQUIRE((codes_s8 == nullptr) == (s8_meta == nullptr), "oq_norm_act_quant_fwd: codes_s8 and s8_meta go together");
  OQ_REQUIRE(!codes_s8 || (ld_s8 % 8 == 0 && (uintptr_t)codes_s8 % 8 == 0), "oq_norm_act_quant_fwd: code rows must be 8-byte aligned");
  OQ_REQUIRE(x && yq && w && rows > 0 && C > 0, "oq_norm_act_quant_fwd: null pointer or bad shape");
  OQ_REQUIRE(kind == 0 || kind == 1, "oq_norm_act_quant_fwd: kind must be 0 (rms) or 1 (layernorm)");
  OQ_REQUIRE(rows < (1L << 31), "oq_norm_act_quant_fwd: too many rows");
  OQ_REQUIRE(n_bits >= 2 && n_bits <= 8, "oq_norm_act_quant_fwd: n_bits %d unsupported", n_bits);
  NormArgs a = {};
  a.x = (const bf16*)x; a.ldx = ldx; a.rows = rows; a.C = C; a.w = w; a.b = b; a.eps = eps; a.kind = kind;
  a.y = (bf16*)y; a.ldy = ldy; a.mean = mean_out; a.rstd = rstd_out;
  a.yq = (bf16*)yq; a.ldyq = ldyq; a.n_bits = n_bits; a.symmetric = symmetric; a.recip_div = recip_div;
  a.s8 = {codes_s8, ld_s8, s8_meta, rows};
  return norm_fwd_launch("oq_norm_act_quant_fwd", a, stream);
}

extern "C" int oq_norm_bwd(const void* x, long rows, int C, long ldx, const float* w, int kind, const float* mean,
                           const float* rstd, const void* dy, long lddy, void* dx, long lddx, const void* dx_add,
                           long ldadd, float* dw_acc, float* db_acc, cudaStream_t stream) {
  OQ_REQUIRE(x && dy && w && rstd && rows > 0 && C > 0, "oq_norm_bwd: null pointer or bad shape");
  OQ_REQUIRE(kind == 0 || (kind == 1 && mean), "oq_norm_bwd: layernorm needs mean");
  OQ_REQUIRE(dx || dw_acc || db_acc, "oq_norm_bwd: no output requested");
  NormArgs a = {};
  a.x = (const bf16*)x; a.ldx = ldx; a.rows = rows; a.C = C; a.w = w; a.kind = kind; a.mean = (float*)mean;
  a.rstd = (float*)rstd; a.dy = (const bf16*)dy; a.lddy = lddy; a.dx = (bf16*)dx; a.lddx = lddx;
  a.dx_add = (const bf16*)dx_add; a.ldadd = ldadd; a.dw_acc = dw_acc; a.db_acc = db_acc;
  const size_t smem = sizeof(float) * C * ((dw_acc ? 1 : 0) + (db_acc ? 1 : 0));
  OQ_REQUIRE(smem <= 200 * 1024, "oq_norm_bwd: C=%d too large for the per-channel accumulators", C);
  const bool vec = C % 8 == 0 && ldx % 8 == 0 && lddy % 8 == 0 && (!dx || lddx % 8 == 0) && (!dx_add || ldadd % 8 == 0) &&
                   (uintptr_t)x % 16 == 0 && (uintptr_t)dy % 16 == 0 && (uintptr_t)dx % 16 == 0 &&
                   (uintptr_t)dx_add % 16 == 0;
  const int grid = (int)std::min<long>(rows, smem ? (long)oq_num_sms() * 2 : rows);
  if (vec && C <= 8192) {
    const bool acc = dw_acc || db_acc;
#define OQ_NB(NV, TPB)                                                     \
  {                                                                        \
    if (acc) norm_bwd_reg<NV, true, TPB><<<grid, TPB, 0, stream>>>(a);     \
    else norm_bwd_reg<NV, false, TPB><<<grid, TPB, 0, stream>>>(a);        \
  }
    if (C <= 2048) OQ_NB(1, 256) else if (C <= 4096) OQ_NB(1, 512) else OQ_NB(2, 512)
#undef OQ_NB
  } else if (vec) {
    if (smem > 40 * 1024) cudaFuncSetAttribute(norm_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    norm_bwd_kernel<true><<<grid, 256, smem, stream>>>(a);
  } else {
    if (smem > 40 * 1024) cudaFuncSetAttribute(norm_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    norm_bwd_kernel<false><<<grid, 256, smem, stream>>>(a);
  }
  return oq_check_launch("oq_norm_bwd");
}
