// Coalesced, vectorised elementwise / reduction kernels of the calibration step:
// SwiGLU / ReLU (models/int_llama_layer.py:44-45, models/int_opt_layer.py:318),
// masked fp32 softmax of QuantMatMul scores (models/int_llama_layer.py:156-160),
// the block-output MSE loss forward+backward (quantize/omniquant.py:142,220-222),
// AdamW on the flat LWC/LET parameter buffer (quantize/omniquant.py:207-208 +
// utils.py:32-46), bias-gradient column sums, truncate_number
// (models/transformation.py:5-20) and a bf16 transpose.  All HBM-bound.
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"

namespace {

// ---------------------------------------------------------------- SwiGLU
// gu: [rows, 2F] = [gate | up];  m = silu(gate) * up
__global__ void __launch_bounds__(256) swiglu_fwd_kernel(const bf16* gu, long ldgu, bf16* m, long ldm, long rows, int F) {
  const long n8 = (long)rows * (F / 8);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    const long r = i / (F / 8);
    const int c = (int)(i % (F / 8)) * 8;
    float g[8], u[8];
    load_bf8(gu + r * ldgu + c, g);
    load_bf8(gu + r * ldgu + F + c, u);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = swiglu_val(g[j], u[j]);
    store_bf8(m + r * ldm + c, g);
  }
}
__global__ void __launch_bounds__(256) swiglu_bwd_kernel(const bf16* gu, long ldgu, const bf16* dm, long lddm, bf16* dgu,
                                                         long lddgu, long rows, int F) {
  const long n8 = (long)rows * (F / 8);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    const long r = i / (F / 8);
    const int c = (int)(i % (F / 8)) * 8;
    float g[8], u[8], d[8];
    load_bf8(gu + r * ldgu + c, g);
    load_bf8(gu + r * ldgu + F + c, u);
    load_bf8(dm + r * lddm + c, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float dg, du;
      swiglu_grad(g[j], u[j], d[j], dg, du);
      g[j] = dg;
      u[j] = du;
    }
    store_bf8(dgu + r * lddgu + c, g);
    store_bf8(dgu + r * lddgu + F + c, u);
  }
}

// ---------------------------------------------------------------- exact (erf) GELU of QuantFalconMLP (nn.GELU(), models/int_falcon_layer.py:25)
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const bf16* x, bf16* y, long n8) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    float v[8];
    load_bf8(x + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.5f * v[j] * (1.0f + erff(v[j] * 0.70710678118654752f));
    store_bf8(y + i * 8, v);
  }
}
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const bf16* x, const bf16* dy, bf16* dx, long n8) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    float v[8], g[8];
    load_bf8(x + i * 8, v);
    load_bf8(dy + i * 8, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float cdf = 0.5f * (1.0f + erff(v[j] * 0.70710678118654752f));
      const float pdf = 0.3989422804014327f * __expf(-0.5f * v[j] * v[j]);
      g[j] *= cdf + v[j] * pdf;
    }
    store_bf8(dx + i * 8, g);
  }
}

// ---------------------------------------------------------------- ReLU (OPT)
__global__ void __launch_bounds__(256) relu_fwd_kernel(const bf16* x, bf16* y, long n8) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    float v[8];
    load_bf8(x + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
    store_bf8(y + i * 8, v);
  }
}
__global__ void __launch_bounds__(256) relu_bwd_kernel(const bf16* x, const bf16* dy, bf16* dx, long n8) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    float v[8], g[8];
    load_bf8(x + i * 8, v);
    load_bf8(dy + i * 8, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = v[j] > 0.0f ? g[j] : 0.0f;
    store_bf8(dx + i * 8, g);
  }
}

// ---------------------------------------------------------------- softmax
// S: fp32 scores [B, Tq, Tk] (already scaled by 1/sqrt(d) in the GEMM epilogue);
// P = softmax(max(S + mask, -FLT_MAX)) in fp32, stored bf16.  causal != 0 masks j > i + (Tk - Tq).
__global__ void __launch_bounds__(256) softmax_fwd_kernel(const float* S, bf16* P, const float* mask, long mask_bstride,
                                                          int Tq, int Tk, int causal) {
  __shared__ float scratch[32];
  const long row = blockIdx.x;  // over B*Tq
  const int i = (int)(row % Tq);
  const long b = row / Tq;
  const float* s = S + row * (long)Tk;
  bf16* p = P + row * (long)Tk;
  const float* mk = mask ? mask + b * mask_bstride + (long)i * Tk : nullptr;
  const int lim = causal ? min(Tk, i + (Tk - Tq) + 1) : Tk;
  float mx[1] = {-INFINITY};
  for (int j = threadIdx.x; j < lim; j += blockDim.x) {
    float v = s[j];
    if (mk) v = fmaxf(v + mk[j], -3.4028234663852886e38f);
    mx[0] = fmaxf(mx[0], v);
  }
  block_reduce<1>(mx, scratch, OpMax(), -INFINITY);
  float sm[1] = {0.f};
  for (int j = threadIdx.x; j < lim; j += blockDim.x) {
    float v = s[j];
    if (mk) v = fmaxf(v + mk[j], -3.4028234663852886e38f);
    sm[0] += __expf(v - mx[0]);
  }
  block_reduce<1>(sm, scratch, OpSum(), 0.0f);
  const float inv = 1.0f / sm[0];
  for (int j = threadIdx.x; j < Tk; j += blockDim.x) {
    float o = 0.0f;
    if (j < lim) {
      float v = s[j];
      if (mk) v = fmaxf(v + mk[j], -3.4028234663852886e38f);
      o = __expf(v - mx[0]) * inv;
    }
    p[j] = f2bf(o);
  }
}
// dS = P * (dP - sum_j P*dP), stored bf16 (may alias P).  causal: row i only holds j <= i + (Tk - Tq);
// the rest of dP was never written by the tile-skipping GEMM and dS is zero there.
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const bf16* P, const float* dP, bf16* dS, int Tq, int Tk, int causal) {
  __shared__ float scratch[32];
  const long row = blockIdx.x;
  const int i = (int)(row % Tq);
  const int lim = causal ? min(Tk, i + (Tk - Tq) + 1) : Tk;
  const bf16* p = P + row * (long)Tk;
  const float* dp = dP + row * (long)Tk;
  bf16* ds = dS + row * (long)Tk;
  float acc[1] = {0.f};
  for (int j = threadIdx.x; j < lim; j += blockDim.x) acc[0] += bf2f(p[j]) * dp[j];
  block_reduce<1>(acc, scratch, OpSum(), 0.0f);
  for (int j = threadIdx.x; j < Tk; j += blockDim.x) {
    float o = 0.0f;
    if (j < lim) o = bf2f(p[j]) * (dp[j] - acc[0]);
    ds[j] = f2bf(o);
  }
}


// 4 scores / gradients from fp32, fp16 or bf16 storage
__device__ __forceinline__ float4 ld4_f32(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4_f16(const __half* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 ld4_bf16(const bf16* p) {
  float t[4];
  load_bf4(p, t);
  return make_float4(t[0], t[1], t[2], t[3]);
}
template <int DT>   // 1 fp32, 2 fp16, 0 bf16
__device__ __forceinline__ float4 ld4_any(const void* base, long idx) {
  if constexpr (DT == 1) return ld4_f32(reinterpret_cast<const float*>(base) + idx);
  else if constexpr (DT == 2) return ld4_f16(reinterpret_cast<const __half*>(base) + idx);
  else return ld4_bf16(reinterpret_cast<const bf16*>(base) + idx);
}

// ---- register-resident softmax rows (Tk <= 1024*NV, Tk % 4 == 0): one read of S / dP, one write of P / dS.
// Causal rows read j <= i only and write zeros only up to the end of the 128-wide diagonal tile: the k-limited
// P.V / dS.K / P^T.dO / dS^T.Q GEMMs (causal modes 2 and 3) never read beyond it.
template <int NV, int DT>
__global__ void __launch_bounds__(256) softmax_fwd_reg(const void* S, bf16* P, const float* mask, long mask_bstride, int Tq,
                                                       int Tk, int causal) {
  __shared__ float scratch[2 * 8];
  const long row = blockIdx.x;
  const int i = (int)(row % Tq);
  const long b = row / Tq;
  const long sbase = row * (long)Tk;
  bf16* p = P + row * (long)Tk;
  const float* mk = mask ? mask + b * mask_bstride + (long)i * Tk : nullptr;
  const int lim = causal ? min(Tk, i + (Tk - Tq) + 1) : Tk;
  const int wlim = causal ? min(Tk, (lim + 127) & ~127) : Tk;
  float4 v[NV];
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int j = 4 * (threadIdx.x + c * 256);
    v[c] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (j < lim) {
      float4 t = ld4_any<DT>(S, sbase + j);
      if (mk) {
        const float4 m4 = *reinterpret_cast<const float4*>(mk + j);
        t.x = fmaxf(t.x + m4.x, -3.4028234663852886e38f); t.y = fmaxf(t.y + m4.y, -3.4028234663852886e38f);
        t.z = fmaxf(t.z + m4.z, -3.4028234663852886e38f); t.w = fmaxf(t.w + m4.w, -3.4028234663852886e38f);
      }
      if (j + 1 >= lim) t.y = -INFINITY;
      if (j + 2 >= lim) t.z = -INFINITY;
      if (j + 3 >= lim) t.w = -INFINITY;
      v[c] = t;
      mx = fmaxf(fmaxf(mx, fmaxf(t.x, t.y)), fmaxf(t.z, t.w));
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  mx = warp_reduce(mx, OpMax());
  if (lane == 0) scratch[warp] = mx;
  __syncthreads();
  mx = scratch[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, scratch[w]);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    v[c].x = __expf(v[c].x - mx); v[c].y = __expf(v[c].y - mx); v[c].z = __expf(v[c].z - mx); v[c].w = __expf(v[c].w - mx);
    sum += (v[c].x + v[c].y) + (v[c].z + v[c].w);
  }
  sum = warp_reduce(sum, OpSum());
  if (lane == 0) scratch[8 + warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += scratch[8 + w];
  const float inv = 1.0f / sum;
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int j = 4 * (threadIdx.x + c * 256);
    if (j < wlim) {
      const float o[4] = {v[c].x * inv, v[c].y * inv, v[c].z * inv, v[c].w * inv};
      store_bf4(p + j, o);
    }
  }
}
template <int NV, int DT>
__global__ void __launch_bounds__(256) softmax_bwd_reg(const bf16* P, const void* dP, bf16* dS, int Tq, int Tk, int causal) {
  __shared__ float scratch[8];
  const long row = blockIdx.x;
  const int i = (int)(row % Tq);
  const int lim = causal ? min(Tk, i + (Tk - Tq) + 1) : Tk;
  const int wlim = causal ? min(Tk, (lim + 127) & ~127) : Tk;
  const bf16* p = P + row * (long)Tk;
  const long dbase = row * (long)Tk;
  bf16* ds = dS + row * (long)Tk;
  float pv[NV][4];
  float4 g[NV];
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int j = 4 * (threadIdx.x + c * 256);
    pv[c][0] = pv[c][1] = pv[c][2] = pv[c][3] = 0.f;
    g[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < lim) {
      load_bf4(p + j, pv[c]);
      g[c] = ld4_any<DT>(dP, dbase + j);
      if (j + 1 >= lim) { pv[c][1] = 0.f; g[c].y = 0.f; }
      if (j + 2 >= lim) { pv[c][2] = 0.f; g[c].z = 0.f; }
      if (j + 3 >= lim) { pv[c][3] = 0.f; g[c].w = 0.f; }
      acc += pv[c][0] * g[c].x + pv[c][1] * g[c].y + pv[c][2] * g[c].z + pv[c][3] * g[c].w;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  acc = warp_reduce(acc, OpSum());
  if (lane == 0) scratch[warp] = acc;
  __syncthreads();
  acc = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) acc += scratch[w];
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int j = 4 * (threadIdx.x + c * 256);
    if (j < wlim) {
      const float o[4] = {pv[c][0] * (g[c].x - acc), pv[c][1] * (g[c].y - acc), pv[c][2] * (g[c].z - acc),
                          pv[c][3] * (g[c].w - acc)};
      store_bf4(ds + j, o);
    }
  }
}

// ---- warp-per-row softmax for 16-bit scores (Tk <= 256*NC, Tk % 8 == 0): 16-byte accesses, NC independent loads in
// flight per lane, no block barrier; causal rows stop at the diagonal as above.
template <int DT>   // 2 fp16, 0 bf16
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  if constexpr (DT == 2) {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(h[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
  } else {
    unpack_bf2(u.x, v[0], v[1]); unpack_bf2(u.y, v[2], v[3]); unpack_bf2(u.z, v[4], v[5]); unpack_bf2(u.w, v[6], v[7]);
  }
}

// Two warps per row, four rows per 256-thread block: a lane holds NC 16-byte chunks (Tk <= 512*NC), the two halves of
// a row meet through shared memory.  Few registers per thread -> 32+ resident warps per SM to cover the HBM latency.
template <int NC, int DT>
__global__ void __launch_bounds__(256, 4) softmax_fwd_warp(const void* S, bf16* P, const float* mask, long mask_bstride,
                                                           long rows, int Tq, int Tk, int causal) {
  __shared__ float red[2][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = warp & 1;
  long row = (long)blockIdx.x * 4 + (warp >> 1);
  const bool live = row < rows;
  if (!live) row = rows - 1;
  const int i = (int)(row % Tq);
  const long b = row / Tq;
  const uint16_t* s = reinterpret_cast<const uint16_t*>(S) + row * (long)Tk;
  bf16* p = P + row * (long)Tk;
  const float* mk = mask ? mask + b * mask_bstride + (long)i * Tk : nullptr;
  const int lim = causal ? min(Tk, i + (Tk - Tq) + 1) : Tk;
  const int wlim = causal ? min(Tk, (lim + 127) & ~127) : Tk;
  const int nc = (lim + 511) >> 9, ncw = (wlim + 511) >> 9;   // 512-column chunks with data / to be written (row-uniform)
  const int j0 = (half * 32 + lane) * 8;
  uint4 raw[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c)
    if (c < nc && c * 512 + j0 < lim) raw[c] = *reinterpret_cast<const uint4*>(s + c * 512 + j0);
  float v[NC][8];
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c < nc) {
      const int j = c * 512 + j0;
      if (j < lim) {
        unpack8<DT>(raw[c], v[c]);
        if (mk) {
          const float4 m0 = *reinterpret_cast<const float4*>(mk + j), m1 = *reinterpret_cast<const float4*>(mk + j + 4);
          const float m[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) v[c][e] = fmaxf(v[c][e] + m[e], -3.4028234663852886e38f);
        }
        if (j + 8 > lim) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (j + e >= lim) v[c][e] = -INFINITY;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) mx = fmaxf(mx, v[c][e]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[c][e] = -INFINITY;
      }
    }
  }
  mx = warp_reduce(mx, OpMax());
  if (lane == 0) red[0][warp] = mx;
  __syncthreads();
  mx = fmaxf(mx, red[0][warp ^ 1]);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c < nc) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[c][e] = __expf(v[c][e] - mx);
      sum += ((v[c][0] + v[c][1]) + (v[c][2] + v[c][3])) + ((v[c][4] + v[c][5]) + (v[c][6] + v[c][7]));
    }
  }
  sum = warp_reduce(sum, OpSum());
  if (lane == 0) red[1][warp] = sum;
  __syncthreads();
  sum = red[1][warp & ~1] + red[1][warp | 1];
  const float inv = 1.0f / sum;
  if (!live) return;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c < ncw) {
      const int j = c * 512 + j0;
      if (j < wlim) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = c < nc ? v[c][e] * inv : 0.f;
        store_bf8(p + j, o);
      }
    }
  }
}

template <int NC>
__global__ void __launch_bounds__(256, 3) softmax_bwd_warp(const bf16* P, const bf16* dP, bf16* dS, long rows, int Tq, int Tk,
                                                           int causal) {
  __shared__ float red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = warp & 1;
  long row = (long)blockIdx.x * 4 + (warp >> 1);
  const bool live = row < rows;
  if (!live) row = rows - 1;
  const int i = (int)(row % Tq);
  const int lim = causal ? min(Tk, i + (Tk - Tq) + 1) : Tk;
  const int wlim = causal ? min(Tk, (lim + 127) & ~127) : Tk;
  const int nc = (lim + 511) >> 9, ncw = (wlim + 511) >> 9;
  const int j0 = (half * 32 + lane) * 8;
  const bf16* p = P + row * (long)Tk;
  const bf16* g = dP + row * (long)Tk;
  bf16* ds = dS + row * (long)Tk;
  uint4 pr[NC], gr[NC];      // the row stays packed in registers between the two passes
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    pr[c] = make_uint4(0u, 0u, 0u, 0u);
    gr[c] = make_uint4(0u, 0u, 0u, 0u);
    if (c < nc && c * 512 + j0 < lim) {
      pr[c] = *reinterpret_cast<const uint4*>(p + c * 512 + j0);
      gr[c] = *reinterpret_cast<const uint4*>(g + c * 512 + j0);
    }
  }
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c < nc) {
      const int j = c * 512 + j0;
      float pv[8], gv[8];
      unpack8<0>(pr[c], pv);
      unpack8<0>(gr[c], gv);
      if (j + 8 > lim) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (j + e >= lim) pv[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) acc += pv[e] * gv[e];
    }
  }
  acc = warp_reduce(acc, OpSum());
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  acc = red[warp & ~1] + red[warp | 1];
  if (!live) return;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c < ncw) {
      const int j = c * 512 + j0;
      if (j < wlim) {
        float pv[8], gv[8];
        unpack8<0>(pr[c], pv);
        unpack8<0>(gr[c], gv);
#pragma unroll
        for (int e = 0; e < 8; ++e) pv[e] = (j + e < lim) ? pv[e] * (gv[e] - acc) : 0.f;
        store_bf8(ds + j, pv);
      }
    }
  }
}

// ---------------------------------------------------------------- MSE
// loss += inv_numel * sum((out-target)^2 [+ (out-target2)^2]);  dout = 2*inv_numel*((out-target) [+ (out-target2)])
__global__ void __launch_bounds__(256) mse_kernel(const bf16* out, const bf16* tgt, const bf16* tgt2, long n8,
                                                  float inv_numel, float* loss_acc, bf16* dout) {
  __shared__ float scratch[32];
  float acc[1] = {0.f};
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    float o[8], t[8], t2[8];
    load_bf8(out + i * 8, o);
    load_bf8(tgt + i * 8, t);
    if (tgt2) load_bf8(tgt2 + i * 8, t2);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float d = o[j] - t[j];
      acc[0] += d * d;
      if (tgt2) {
        const float d2 = o[j] - t2[j];
        acc[0] += d2 * d2;
        d += d2;
      }
      o[j] = 2.0f * inv_numel * d;
    }
    if (dout) store_bf8(dout + i * 8, o);
  }
  block_reduce<1>(acc, scratch, OpSum(), 0.0f);
  if (threadIdx.x == 0 && loss_acc) atomicAdd(loss_acc, acc[0] * inv_numel);
}

// ---------------------------------------------------------------- AdamW
// torch.optim.AdamW (amsgrad=False, maximize=False) on a flat fp32 segment.
__global__ void __launch_bounds__(256) adamw_kernel(float* p, const float* g, float* m, float* v, long n, float lr,
                                                    float b1, float b2, float eps, float wd, float grad_scale,
                                                    const int* step_ptr) {
  const int step = *step_ptr;
  const float bc1 = 1.0f - powf(b1, (float)step);
  const float bc2 = 1.0f - powf(b2, (float)step);
  const float step_size = lr / bc1;
  const float bc2_sqrt = sqrtf(bc2);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    float pi = p[i];
    pi *= (1.0f - lr * wd);
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - step_size * (mi / denom);
  }
}
__global__ void step_inc_kernel(int* step) { *step += 1; }

__global__ void __launch_bounds__(256) truncate_kernel(float* p, long n, float thr) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    p[i] = oq_truncate(p[i], thr);
}

// ---------------------------------------------------------------- column sums
// out[c] (+)= sum_r X[r, c]; X bf16 [rows, C].  grid (C/256 tiles, row chunks).
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* X, long ld, long rows, int C, float* out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const long r0 = (long)blockIdx.y * rows / gridDim.y, r1 = (long)(blockIdx.y + 1) * rows / gridDim.y;
  float acc = 0.f;
  for (long r = r0; r < r1; ++r) acc += bf2f(X[r * ld + c]);
  atomicAdd(out + c, acc);
}

// V = 2 / 4 / 8 columns per thread (4 / 8 / 16-byte loads), four rows in flight; at most 32 row chunks, because every chunk ends
// in one atomicAdd per column and those serialise per address (256 chunks: 26 us whatever the size).  The scalar kernel above
// moved 2 bytes per load instruction: 3.3-3.6 TB/s on the [2048, 4096..22016] gradients of a LLaMA-7B step.
template <int V>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const bf16* X, long ld, long rows, int C, float* out) {
  const int c = (blockIdx.x * 256 + threadIdx.x) * V;
  if (c >= C) return;
  const long r0 = (long)blockIdx.y * rows / gridDim.y, r1 = (long)(blockIdx.y + 1) * rows / gridDim.y;
  float acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.f;
  const bf16* px = X + c;
  auto load = [&](long r, float (&v)[V]) {
    if constexpr (V == 8) {
      load_bf8(px + r * ld, v);
    } else if constexpr (V == 4) {
      load_bf4(px + r * ld, v);
    } else {
      unpack_bf2(*reinterpret_cast<const uint32_t*>(px + r * ld), v[0], v[1]);
    }
  };
  long r = r0;
  for (; r + 4 <= r1; r += 4) {
    float v[4][V];
#pragma unroll
    for (int u = 0; u < 4; ++u) load(r + u, v[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] += v[u][j];
  }
  for (; r < r1; ++r) {
    float v[V];
    load(r, v);
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] += v[j];
  }
#pragma unroll
  for (int j = 0; j < V; ++j) atomicAdd(out + c + j, acc[j]);
}

// ---------------------------------------------------------------- transpose
__global__ void transpose_kernel(const bf16* in, long ldi, bf16* out, long ldo, int R, int C) {
  __shared__ bf16 tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[j][threadIdx.x] = in[(long)r * ldi + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < C) out[(long)c * ldo + r] = tile[threadIdx.x][j];
  }
}


// ---------------------------------------------------------------- LET vector folds (O(H) work, one launch each)
// Packed layouts (H = hidden size, F1 = rows of the first MLP weight group):
//   let   [7H]      qkv_s | out_s | fc1_s | qkt_s | qkv_d | out_d | fc1_d        (parameters or their gradients)
//   mv    [4H+F1]   W_q@d | W_k@d | W_v@d | W_o@d_out | W_fc1@d_fc1                (from oq_weight_quant_fwd)
//   tmpv  [9H+F1]   ln1.w' | ln1.b' | ln2.w' | ln2.b' | b_q' b_k' b_v' | b_o' | b_fc1' | b_q'*qscale
//   acc   [10H]     dcol_qkv | dcol_out | dcol_fc1 | dsh_qkv | dsh_out | dsh_fc1 | dw1 | db1 | dw2 | db2
// Same op order as models/transformation.py:28-33,37-40,49-50,56-59,66-69.
__global__ void __launch_bounds__(256) let_fold_vec_kernel(int H, int F1, const float* let, const float* ln1w, const float* ln1b,
                                                           const float* ln2w, const float* ln2b, const float* rb_qkv,
                                                           const float* rb_o, const float* rb_fc1, const float* mv,
                                                           float qscale, float* tmpv, int part) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float *s1 = let, *so = let + H, *s2 = let + 2 * H, *sk = let + 3 * H, *d1 = let + 4 * H, *dout = let + 5 * H,
              *d2 = let + 6 * H;
  if (i < H && part != 2) {
    tmpv[i] = OQ_DIV(ln1w[i], s1[i]);
    tmpv[H + i] = ln1b ? OQ_DIV(OQ_SUB(ln1b[i], d1[i]), s1[i]) : OQ_DIV(OQ_MUL(-1.0f, d1[i]), s1[i]);
    tmpv[2 * H + i] = OQ_DIV(ln2w[i], s2[i]);
    tmpv[3 * H + i] = ln2b ? OQ_DIV(OQ_SUB(ln2b[i], d2[i]), s2[i]) : OQ_DIV(OQ_MUL(-1.0f, d2[i]), s2[i]);
    const float bq = rb_qkv ? OQ_ADD(rb_qkv[i], mv[i]) : mv[i];
    const float bk = rb_qkv ? OQ_ADD(rb_qkv[H + i], mv[H + i]) : mv[H + i];
    const float bv = rb_qkv ? OQ_ADD(rb_qkv[2 * H + i], mv[2 * H + i]) : mv[2 * H + i];
    const float bqt = OQ_DIV(bq, sk[i]);
    tmpv[4 * H + i] = bqt;
    tmpv[5 * H + i] = OQ_MUL(bk, sk[i]);
    tmpv[6 * H + i] = OQ_DIV(OQ_SUB(bv, dout[i]), so[i]);
    tmpv[8 * H + F1 + i] = bqt * qscale;
  }
  if (i < H && part != 1) tmpv[7 * H + i] = rb_o ? OQ_ADD(rb_o[i], mv[3 * H + i]) : mv[3 * H + i];
  if (i < F1 && part != 1) tmpv[8 * H + i] = rb_fc1 ? OQ_ADD(rb_fc1[i], mv[4 * H + i]) : mv[4 * H + i];
}
// e_q = db_q*qscale/s_qkt, e_k = db_k*s_qkt, e_v = db_v/s_out   (dL/d(W@shift) per output row)
__global__ void __launch_bounds__(256) let_prep_bwd_kernel(int H, const float* let, const float* dbqkv, float qscale, float* e) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H) return;
  const float so = let[H + i], sk = let[3 * H + i];
  e[i] = dbqkv[i] * qscale / sk;
  e[H + i] = dbqkv[H + i] * sk;
  e[2 * H + i] = dbqkv[2 * H + i] / so;
}
// LET gradient assembly (DESIGN.md / SURVEY.md section 3.4b)
__global__ void __launch_bounds__(256) let_grad_vec_kernel(int H, int F1, const float* let, const float* tmpv, const float* acc,
                                                           const float* drow, const float* dbqkv, float qscale, float* g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H) return;
  const float s1 = let[i], so = let[H + i], s2 = let[2 * H + i], sk = let[3 * H + i];
  const float ln1w = tmpv[i], ln1b = tmpv[H + i], ln2w = tmpv[2 * H + i], ln2b = tmpv[3 * H + i];
  const float bq = tmpv[4 * H + i], bk = tmpv[5 * H + i], bv = tmpv[6 * H + i];
  const float dcol_qkv = acc[i], dcol_out = acc[H + i], dcol_fc1 = acc[2 * H + i], dsh_qkv = acc[3 * H + i],
              dsh_out = acc[4 * H + i], dsh_fc1 = acc[5 * H + i], dw1 = acc[6 * H + i], db1 = acc[7 * H + i],
              dw2 = acc[8 * H + i], db2 = acc[9 * H + i];
  const float dbq = dbqkv[i] * qscale, dbk = dbqkv[H + i], dbv = dbqkv[2 * H + i];
  g[i] = dcol_qkv - (dw1 * ln1w + db1 * ln1b) / s1;                                     // qkv_s
  g[H + i] = dcol_out - (drow[2 * H + i] / so + dbv * bv) / so;                           // out_s
  g[2 * H + i] = dcol_fc1 - (dw2 * ln2w + db2 * ln2b) / s2;                               // fc1_s
  g[3 * H + i] = (-drow[i] / sk - dbq * bq) / sk + (drow[H + i] + dbk * bk / sk);         // qkt_s
  g[4 * H + i] = dsh_qkv - db1 / s1;                                                      // qkv_d
  g[5 * H + i] = dsh_out - dbv / so;                                                      // out_d
  g[6 * H + i] = dsh_fc1 - db2 / s2;                                                      // fc1_d
}

int grid_for(long n, int per_thread = 1) {
  long blocks = (n + 256L * per_thread - 1) / (256L * per_thread);
  long cap = (long)oq_num_sms() * 8;
  return (int)std::max<long>(1, std::min(blocks, cap));
}
bool al16(const void* p) { return ((uintptr_t)p % 16) == 0; }

}  // namespace

extern "C" int oq_swiglu_fwd(const void* gu, long ldgu, void* m, long ldm, long rows, int F, cudaStream_t stream) {
  OQ_REQUIRE(gu && m && rows > 0 && F > 0 && F % 8 == 0 && ldgu % 8 == 0 && ldm % 8 == 0 && al16(gu) && al16(m),
             "oq_swiglu_fwd: need F, strides multiples of 8 and 16-byte aligned pointers");
  swiglu_fwd_kernel<<<grid_for(rows * (F / 8)), 256, 0, stream>>>((const bf16*)gu, ldgu, (bf16*)m, ldm, rows, F);
  return oq_check_launch("oq_swiglu_fwd");
}
extern "C" int oq_swiglu_bwd(const void* gu, long ldgu, const void* dm, long lddm, void* dgu, long lddgu, long rows,
                             int F, cudaStream_t stream) {
  OQ_REQUIRE(gu && dm && dgu && rows > 0 && F > 0 && F % 8 == 0 && ldgu % 8 == 0 && lddm % 8 == 0 && lddgu % 8 == 0 &&
                 al16(gu) && al16(dm) && al16(dgu),
             "oq_swiglu_bwd: need F, strides multiples of 8 and 16-byte aligned pointers");
  swiglu_bwd_kernel<<<grid_for(rows * (F / 8)), 256, 0, stream>>>((const bf16*)gu, ldgu, (const bf16*)dm, lddm,
                                                                  (bf16*)dgu, lddgu, rows, F);
  return oq_check_launch("oq_swiglu_bwd");
}
extern "C" int oq_relu_fwd(const void* x, void* y, long n, cudaStream_t stream) {
  OQ_REQUIRE(x && y && n > 0 && n % 8 == 0 && al16(x) && al16(y), "oq_relu_fwd: n %% 8 and 16-byte alignment required");
  relu_fwd_kernel<<<grid_for(n / 8), 256, 0, stream>>>((const bf16*)x, (bf16*)y, n / 8);
  return oq_check_launch("oq_relu_fwd");
}
extern "C" int oq_relu_bwd(const void* x, const void* dy, void* dx, long n, cudaStream_t stream) {
  OQ_REQUIRE(x && dy && dx && n > 0 && n % 8 == 0 && al16(x) && al16(dy) && al16(dx),
             "oq_relu_bwd: n %% 8 and 16-byte alignment required");
  relu_bwd_kernel<<<grid_for(n / 8), 256, 0, stream>>>((const bf16*)x, (const bf16*)dy, (bf16*)dx, n / 8);
  return oq_check_launch("oq_relu_bwd");
}
extern "C" int oq_softmax_fwd(const void* S, int s_dtype, void* P, const float* mask, long mask_bstride, long batch, int Tq,
                              int Tk, int causal, cudaStream_t stream) {
  OQ_REQUIRE(S && P && batch > 0 && Tq > 0 && Tk > 0 && batch * Tq < (1L << 31), "oq_softmax_fwd: bad arguments");
  OQ_REQUIRE(s_dtype == 1 || s_dtype == 2, "oq_softmax_fwd: scores must be fp32 (1) or fp16 (2)");
  const unsigned grid = (unsigned)(batch * Tq);
  const bool fast = Tk % 4 == 0 && al16(S) && ((uintptr_t)P % 8 == 0) && al16(mask) && Tk <= 8192;
  OQ_REQUIRE(fast || s_dtype == 1, "oq_softmax_fwd: fp16 scores need Tk %% 4 == 0, Tk <= 8192 and aligned pointers");
#define OQ_SMF(NV)                                                                                                   \
  {                                                                                                                  \
    if (s_dtype == 1) softmax_fwd_reg<NV, 1><<<grid, 256, 0, stream>>>(S, (bf16*)P, mask, mask_bstride, Tq, Tk, causal); \
    else softmax_fwd_reg<NV, 2><<<grid, 256, 0, stream>>>(S, (bf16*)P, mask, mask_bstride, Tq, Tk, causal);          \
  }
  const bool wrow = s_dtype == 2 && Tk % 8 == 0 && Tk <= 2048 && al16(S) && al16(P) && al16(mask);
  const long rows = batch * Tq;
  const unsigned wgrid = (unsigned)((rows + 3) / 4);
  if (wrow && Tk <= 1024) softmax_fwd_warp<2, 2><<<wgrid, 256, 0, stream>>>(S, (bf16*)P, mask, mask_bstride, rows, Tq, Tk, causal);
  else if (wrow) softmax_fwd_warp<4, 2><<<wgrid, 256, 0, stream>>>(S, (bf16*)P, mask, mask_bstride, rows, Tq, Tk, causal);
  else if (fast && Tk <= 1024) OQ_SMF(1)
  else if (fast && Tk <= 2048) OQ_SMF(2)
  else if (fast && Tk <= 4096) OQ_SMF(4)
  else if (fast) OQ_SMF(8)
  else softmax_fwd_kernel<<<grid, 256, 0, stream>>>((const float*)S, (bf16*)P, mask, mask_bstride, Tq, Tk, causal);
#undef OQ_SMF
  return oq_check_launch("oq_softmax_fwd");
}
extern "C" int oq_softmax_bwd(const void* P, const void* dP, int dp_dtype, void* dS, long batch, int Tq, int Tk, int causal,
                              cudaStream_t stream) {
  OQ_REQUIRE(P && dP && dS && batch > 0 && Tq > 0 && Tk > 0 && batch * Tq < (1L << 31), "oq_softmax_bwd: bad arguments");
  OQ_REQUIRE(dp_dtype == 0 || dp_dtype == 1, "oq_softmax_bwd: dP must be bf16 (0) or fp32 (1)");
  const unsigned grid = (unsigned)(batch * Tq);
  const bool fast = Tk % 4 == 0 && al16(dP) && ((uintptr_t)P % 8 == 0) && ((uintptr_t)dS % 8 == 0) && Tk <= 8192;
  OQ_REQUIRE(fast || dp_dtype == 1, "oq_softmax_bwd: bf16 dP needs Tk %% 4 == 0, Tk <= 8192 and aligned pointers");
#define OQ_SMB(NV)                                                                                                  \
  {                                                                                                                 \
    if (dp_dtype == 1) softmax_bwd_reg<NV, 1><<<grid, 256, 0, stream>>>((const bf16*)P, dP, (bf16*)dS, Tq, Tk, causal); \
    else softmax_bwd_reg<NV, 0><<<grid, 256, 0, stream>>>((const bf16*)P, dP, (bf16*)dS, Tq, Tk, causal);           \
  }
  const bool wrow = dp_dtype == 0 && Tk % 8 == 0 && Tk <= 2048 && al16(dP) && al16(P) && al16(dS);
  const long rows = batch * Tq;
  const unsigned wgrid = (unsigned)((rows + 3) / 4);
  if (wrow && Tk <= 1024) softmax_bwd_warp<2><<<wgrid, 256, 0, stream>>>((const bf16*)P, (const bf16*)dP, (bf16*)dS, rows, Tq, Tk, causal);
  else if (wrow) softmax_bwd_warp<4><<<wgrid, 256, 0, stream>>>((const bf16*)P, (const bf16*)dP, (bf16*)dS, rows, Tq, Tk, causal);
  else if (fast && Tk <= 1024) OQ_SMB(1)
  else if (fast && Tk <= 2048) OQ_SMB(2)
  else if (fast && Tk <= 4096) OQ_SMB(4)
  else if (fast) OQ_SMB(8)
  else softmax_bwd_kernel<<<grid, 256, 0, stream>>>((const bf16*)P, (const float*)dP, (bf16*)dS, Tq, Tk, causal);
#undef OQ_SMB
  return oq_check_launch("oq_softmax_bwd");
}
extern "C" int oq_mse_fwd_bwd(const void* out, const void* target, const void* target2, long n, float inv_numel,
                              float* loss_acc, void* dout, cudaStream_t stream) {
  OQ_REQUIRE(out && target && n > 0 && n % 8 == 0 && al16(out) && al16(target) && al16(target2) && al16(dout),
             "oq_mse_fwd_bwd: n %% 8 and 16-byte alignment required");
  mse_kernel<<<grid_for(n / 8, 2), 256, 0, stream>>>((const bf16*)out, (const bf16*)target, (const bf16*)target2, n / 8,
                                                     inv_numel, loss_acc, (bf16*)dout);
  return oq_check_launch("oq_mse_fwd_bwd");
}
extern "C" int oq_adamw(float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2,
                        float eps, float weight_decay, float grad_scale, const int* step_dev, cudaStream_t stream) {
  OQ_REQUIRE(p && g && m && v && step_dev && n > 0, "oq_adamw: null pointer or empty segment");
  adamw_kernel<<<grid_for(n), 256, 0, stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, grad_scale, step_dev);
  return oq_check_launch("oq_adamw");
}
extern "C" int oq_step_inc(int* step_dev, cudaStream_t stream) {
  OQ_REQUIRE(step_dev, "oq_step_inc: null pointer");
  step_inc_kernel<<<1, 1, 0, stream>>>(step_dev);
  return oq_check_launch("oq_step_inc");
}
extern "C" int oq_truncate(float* p, long n, float threshold, cudaStream_t stream) {
  OQ_REQUIRE(p && n > 0, "oq_truncate: null pointer or empty");
  truncate_kernel<<<grid_for(n), 256, 0, stream>>>(p, n, threshold);
  return oq_check_launch("oq_truncate");
}
extern "C" int oq_colsum(const void* X, long ld, long rows, int C, float* out, int accumulate, cudaStream_t stream) {
  OQ_REQUIRE(X && out && rows > 0 && C > 0, "oq_colsum: bad arguments");
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * C, stream);
    if (e != cudaSuccess) return (int)e;
  }
  if (C % 8 == 0 && ld % 8 == 0 && (uintptr_t)X % 16 == 0) {
    const unsigned gy = (unsigned)std::max<long>(1, std::min<long>(rows / 8, 32));
    // the widest loads that still give every SM two blocks
    const int V = ((long)(C / 8 + 255) / 256 * gy >= 2L * oq_num_sms()) ? 8 : ((long)(C / 4 + 255) / 256 * gy >= 2L * oq_num_sms()) ? 4 : 2;
    dim3 grid((C / V + 255) / 256, gy);
    if (V == 8) colsum_vec_kernel<8><<<grid, 256, 0, stream>>>((const bf16*)X, ld, rows, C, out);
    else if (V == 4) colsum_vec_kernel<4><<<grid, 256, 0, stream>>>((const bf16*)X, ld, rows, C, out);
    else colsum_vec_kernel<2><<<grid, 256, 0, stream>>>((const bf16*)X, ld, rows, C, out);
    return oq_check_launch("oq_colsum");
  }
  dim3 grid((C + 255) / 256, (unsigned)std::max<long>(1, std::min<long>(rows / 32, 64)));
  colsum_kernel<<<grid, 256, 0, stream>>>((const bf16*)X, ld, rows, C, out);
  return oq_check_launch("oq_colsum");
}
extern "C" int oq_transpose_bf16(const void* in, long ldi, void* out, long ldo, int R, int C, cudaStream_t stream) {
  OQ_REQUIRE(in && out && R > 0 && C > 0, "oq_transpose_bf16: bad arguments");
  dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
  transpose_kernel<<<grid, block, 0, stream>>>((const bf16*)in, ldi, (bf16*)out, ldo, R, C);
  return oq_check_launch("oq_transpose_bf16");
}

extern "C" int oq_let_fold_vec(int H, int F1, const float* let_params, const float* ln1w, const float* ln1b, const float* ln2w,
                               const float* ln2b, const float* rawb_qkv, const float* rawb_o, const float* rawb_fc1,
                               const float* mv, float q_bias_scale, float* tmpv, int part, cudaStream_t stream) {
  OQ_REQUIRE(H > 0 && F1 > 0 && let_params && ln1w && ln2w && mv && tmpv, "oq_let_fold_vec: bad arguments");
  OQ_REQUIRE(part >= 0 && part <= 2, "oq_let_fold_vec: part must be 0 (all), 1 (norms + q/k/v biases) or 2 (o / fc1 biases)");
  let_fold_vec_kernel<<<(std::max(H, F1) + 255) / 256, 256, 0, stream>>>(H, F1, let_params, ln1w, ln1b, ln2w, ln2b, rawb_qkv,
                                                                          rawb_o, rawb_fc1, mv, q_bias_scale, tmpv, part);
  return oq_check_launch("oq_let_fold_vec");
}
extern "C" int oq_let_prep_bwd(int H, const float* let_params, const float* dbqkv, float q_bias_scale, float* e,
                               cudaStream_t stream) {
  OQ_REQUIRE(H > 0 && let_params && dbqkv && e, "oq_let_prep_bwd: bad arguments");
  let_prep_bwd_kernel<<<(H + 255) / 256, 256, 0, stream>>>(H, let_params, dbqkv, q_bias_scale, e);
  return oq_check_launch("oq_let_prep_bwd");
}
extern "C" int oq_let_grad_vec(int H, int F1, const float* let_params, const float* tmpv, const float* acc, const float* drow,
                               const float* dbqkv, float q_bias_scale, float* let_grads, cudaStream_t stream) {
  OQ_REQUIRE(H > 0 && let_params && tmpv && acc && drow && dbqkv && let_grads, "oq_let_grad_vec: bad arguments");
  let_grad_vec_kernel<<<(H + 255) / 256, 256, 0, stream>>>(H, F1, let_params, tmpv, acc, drow, dbqkv, q_bias_scale, let_grads);
  return oq_check_launch("oq_let_grad_vec");
}

extern "C" int oq_gelu_fwd(const void* x, void* y, long n, cudaStream_t stream) {
  OQ_REQUIRE(x && y && n > 0 && n % 8 == 0 && al16(x) && al16(y), "oq_gelu_fwd: n %% 8 and 16-byte alignment required");
  gelu_fwd_kernel<<<grid_for(n / 8), 256, 0, stream>>>((const bf16*)x, (bf16*)y, n / 8);
  return oq_check_launch("oq_gelu_fwd");
}
extern "C" int oq_gelu_bwd(const void* x, const void* dy, void* dx, long n, cudaStream_t stream) {
  OQ_REQUIRE(x && dy && dx && n > 0 && n % 8 == 0 && al16(x) && al16(dy) && al16(dx),
             "oq_gelu_bwd: n %% 8 and 16-byte alignment required");
  gelu_bwd_kernel<<<grid_for(n / 8), 256, 0, stream>>>((const bf16*)x, (const bf16*)dy, (bf16*)dx, n / 8);
  return oq_check_launch("oq_gelu_bwd");
}
