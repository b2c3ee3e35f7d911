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

}  // namespace

extern "C" int oq_norm_fwd(const void* x, long rows, int C, long ldx, const float* w, const float* b, float eps,
                           int kind, void* y, long ldy, float* mean_out, float* rstd_out, cudaStream_t stream) {
  OQ_REQUIRE(x && y && w && rows > 0 && C > 0, "oq_norm_fwd: null pointer or bad shape");
  OQ_REQUIRE(kind == 0 || kind == 1, "oq_norm_fwd: kind must be 0 (rms) or 1 (layernorm)");
  OQ_REQUIRE(rows < (1L << 31), "oq_norm_fwd: too many rows");
  NormArgs a = {};
  a.x = (const bf16*)x; a.ldx = ldx; a.rows = rows; a.C = C; a.w = w; a.b = b; a.eps = eps; a.kind = kind;
  a.y = (bf16*)y; a.ldy = ldy; a.mean = mean_out; a.rstd = rstd_out;
  const bool vec = C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0;
  if (vec) norm_fwd_kernel<true><<<(unsigned)rows, 256, 0, stream>>>(a);
  else norm_fwd_kernel<false><<<(unsigned)rows, 256, 0, stream>>>(a);
  return oq_check_launch("oq_norm_fwd");
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
  if (vec) {
    if (smem > 40 * 1024) cudaFuncSetAttribute(norm_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    norm_bwd_kernel<true><<<grid, 256, smem, stream>>>(a);
  } else {
    if (smem > 40 * 1024) cudaFuncSetAttribute(norm_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    norm_bwd_kernel<false><<<grid, 256, smem, stream>>>(a);
  }
  return oq_check_launch("oq_norm_bwd");
}
