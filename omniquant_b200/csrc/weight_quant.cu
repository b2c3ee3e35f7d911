// K0: LET fold + LWC clip + fake-quant of a weight matrix, and its backward.
//
// Replaces, for one QuantLinear weight, the chain the reference runs as ~20-40
// separate ATen launches per step:
//   models/transformation.py:35-41,49-60,66-69   (W*s_col {*,/} s_row, W@shift)
//   quantize/quantizer.py:122-147                (group min/max, sigmoid clip, scale, zero point)
//   quantize/quantizer.py:84-105                 (round / clamp / dequant)
// and in the backward the autograd of all of the above.
//
// HBM-bound: forward reads 4 B/weight (fp32 master) and writes 2 B (bf16
// operand for the GEMM); backward reads 4 B + 2 B (bf16 dW_q).  The second pass
// over a row/group re-reads it from L1/L2 (a row is <= 44 KB at LLaMA-7B).
//
// Team = one warp per quantisation group (group_size > 0) or one 256-thread
// block per output channel (per-channel).  Grid = a multiple of the SM count,
// teams stride over groups/rows.
#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace {

struct WqArgs {
  const float* W;
  int N, K;
  const float* col;  // [K] or null
  const float* row;  // [N] or null
  int row_op;        // 0 none 1 mul 2 div
  // optional row segments (one launch over several stacked weights, e.g. q | k | v): rows [seg_end[s-1], seg_end[s]) use
  // seg_row[s] (indexed from the segment's first row) with seg_op[s]; nseg == 0: `row` / `row_op` apply to every row
  int nseg; int seg_end[3]; const float* seg_row[3]; int seg_op[3];
  const float* shift;  // [K] or null
  float* mv;           // [N] W@shift
  int n_bits, symmetric, group, recip_div;
  const float* up;   // [N*G] or null
  const float* low;
  // forward outputs
  bf16* q_bf;  long ld_bf;
  float* q_f32;
  uint8_t* codes;
  int8_t* codes8; float* s8_t; float* s8_beta;   // exact-integer side output (per-channel fast path): code - z0, t, beta
  float* scale_out; float* zp_out; float* xmin_out; float* xmax_out;
  // backward inputs
  const void* g; int g_is_f32; long ld_g;
  const float* xmin_in; const float* xmax_in;
  const float* scale_in; const float* zp_in;  // optional: scale / zero point saved by the forward
  const float* e_row;  // [N] or null
  // backward outputs
  float* d_up; float* d_low;   // [N*G] written
  float* dcol;    // [K] atomically accumulated
  float* drow;    // [N] (per-channel: written; grouped: atomically accumulated)
  float* dshift;  // [K] atomically accumulated
  float* dW;      // [N,K] full gradient w.r.t. the folded weight (optional)
};

template <bool TEAM_WARP>
struct Team {
  __device__ static int size() { return TEAM_WARP ? 32 : blockDim.x; }
  __device__ static int rank() { return TEAM_WARP ? (threadIdx.x & 31) : threadIdx.x; }
  template <int NV, typename Op>
  __device__ static void reduce(float (&v)[NV], float* scratch, Op op, float ident) {
    if (TEAM_WARP) {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = warp_reduce(v[i], op);
    } else {
      block_reduce<NV>(v, scratch, op, ident);
    }
  }
};

// row scale and fold operation of row n (segment table or the launch-wide vector); op 0: rs = 1
__device__ __forceinline__ void wq_row_fold(const WqArgs& a, long n, float& rs, int& op) {
  if (a.nseg == 0) {
    op = a.row ? a.row_op : 0;
    rs = op ? a.row[n] : 1.0f;
    return;
  }
  int s = 0;
  while (s + 1 < a.nseg && n >= a.seg_end[s]) ++s;
  const long first = s ? a.seg_end[s - 1] : 0;
  op = a.seg_row[s] ? a.seg_op[s] : 0;
  rs = op ? a.seg_row[s][n - first] : 1.0f;
}

__device__ __forceinline__ float load_g(const WqArgs& a, long idx) {
  return a.g_is_f32 ? reinterpret_cast<const float*>(a.g)[idx] : bf2f(reinterpret_cast<const bf16*>(a.g)[idx]);
}

// ------------------------------------------------------------------ forward
template <bool TEAM_WARP, bool VEC>
__global__ void __launch_bounds__(256) wq_fwd_kernel(WqArgs a) {
  __shared__ float scratch[3 * 32];
  const int G = a.group ? (a.K + a.group - 1) / a.group : 1;
  const int gsz = a.group ? a.group : a.K;
  const long n_items = (long)a.N * G;
  const int tsz = Team<TEAM_WARP>::size(), tr = Team<TEAM_WARP>::rank();
  long item = TEAM_WARP ? ((long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) : blockIdx.x;
  const long stride = TEAM_WARP ? (long)gridDim.x * (blockDim.x >> 5) : gridDim.x;
  const bool has_col = a.col != nullptr;
  const int lwc = a.up != nullptr;
  constexpr int V = VEC ? 4 : 1;

  for (; item < n_items; item += stride) {
    const int n = (int)(item / G), gi = (int)(item % G);
    const int k0 = gi * gsz, k1 = min(k0 + gsz, a.K);
    const float* Wr = a.W + (long)n * a.K;
    const float rs = a.row ? a.row[n] : 1.0f;
    const int rop = a.row ? a.row_op : 0;
    float red[3] = {INFINITY, -INFINITY, 0.0f};  // min, max, matvec
    for (int k = k0 + tr * V; k < k1; k += tsz * V) {
      float w[V];
      if constexpr (VEC) {
        float4 t = *reinterpret_cast<const float4*>(Wr + k);
        w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
      } else {
        w[0] = Wr[k];
      }
#pragma unroll
      for (int j = 0; j < V; ++j) {
        if (a.shift) red[2] += w[j] * a.shift[k + j];
        float f = oq_fold(w[j], has_col ? a.col[k + j] : 1.0f, has_col, rs, rop);
        red[0] = fminf(red[0], f);
        red[1] = fmaxf(red[1], f);
      }
    }
    if (k0 + gsz > a.K) {  // zero padded tail group (quantizer.py:126-129)
      red[0] = fminf(red[0], 0.0f);
      red[1] = fmaxf(red[1], 0.0f);
    }
    {
      float mm[2] = {-red[0], red[1]}, sm[1] = {red[2]};
      Team<TEAM_WARP>::reduce(mm, scratch, OpMax(), -INFINITY);
      if (a.shift) Team<TEAM_WARP>::reduce(sm, scratch, OpSum(), 0.0f);
      red[0] = -mm[0]; red[1] = mm[1]; red[2] = sm[0];
    }
    const bool ident = a.n_bits >= 16;   // fold only: UniformAffineQuantizer.forward returns x (quantizer.py:109)
    const float su = lwc ? oq_sigmoidf(a.up[item]) : 1.0f;
    const float sl = lwc ? oq_sigmoidf(a.low[item]) : 1.0f;
    OqQParams q = {1.0f, 0.0f, 0.0f};
    if (!ident) q = oq_qparams(red[0], red[1], su, sl, lwc, a.n_bits, a.symmetric, a.recip_div);
    if (tr == 0) {
      if (a.scale_out) a.scale_out[item] = q.scale;
      if (a.zp_out) a.zp_out[item] = q.zp;
      if (a.xmin_out) a.xmin_out[item] = red[0];
      if (a.xmax_out) a.xmax_out[item] = red[1];
      if (a.shift && a.mv) {
        if (G == 1) a.mv[n] = red[2];
        else atomicAdd(a.mv + n, red[2]);
      }
    }
    for (int k = k0 + tr * V; k < k1; k += tsz * V) {
      float w[V], y[V], c[V];
      if constexpr (VEC) {
        float4 t = *reinterpret_cast<const float4*>(Wr + k);
        w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
      } else {
        w[0] = Wr[k];
      }
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float f = oq_fold(w[j], has_col ? a.col[k + j] : 1.0f, has_col, rs, rop);
        c[j] = ident ? 0.0f : oq_code(f, q);
        y[j] = ident ? f : oq_dequant(c[j], q);
      }
      const long o = (long)n * a.K + k;
      if constexpr (VEC) {
        if (a.q_bf) store_bf4(a.q_bf + (long)n * a.ld_bf + k, y);
        if (a.q_f32) *reinterpret_cast<float4*>(a.q_f32 + o) = make_float4(y[0], y[1], y[2], y[3]);
        if (a.codes)
          *reinterpret_cast<uchar4*>(a.codes + o) =
              make_uchar4((unsigned char)c[0], (unsigned char)c[1], (unsigned char)c[2], (unsigned char)c[3]);
      } else {
        if (a.q_bf) a.q_bf[(long)n * a.ld_bf + k] = f2bf(y[0]);
        if (a.q_f32) a.q_f32[o] = y[0];
        if (a.codes) a.codes[o] = (unsigned char)c[0];
      }
    }
  }
}

// ------------------------------------------------------------------ backward
// colacc/shacc: dynamic shared [K] accumulators when dcol / dshift are wanted.
template <bool TEAM_WARP>
__global__ void __launch_bounds__(256) wq_bwd_kernel(WqArgs a) {
  extern __shared__ float dsm[];
  __shared__ float scratch[4 * 32];
  float* colacc = a.dcol ? dsm : nullptr;
  float* shacc = a.dshift ? (dsm + (a.dcol ? a.K : 0)) : nullptr;
  for (int k = threadIdx.x; k < a.K; k += blockDim.x) {
    if (colacc) colacc[k] = 0.0f;
    if (shacc) shacc[k] = 0.0f;
  }
  __syncthreads();

  const int G = a.group ? (a.K + a.group - 1) / a.group : 1;
  const int gsz = a.group ? a.group : a.K;
  const long n_items = (long)a.N * G;
  const int tsz = Team<TEAM_WARP>::size(), tr = Team<TEAM_WARP>::rank();
  long item = TEAM_WARP ? ((long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) : blockIdx.x;
  const long stride = TEAM_WARP ? (long)gridDim.x * (blockDim.x >> 5) : gridDim.x;
  const bool has_col = a.col != nullptr;
  const int lwc = a.up != nullptr;

  for (; item < n_items; item += stride) {
    const int n = (int)(item / G), gi = (int)(item % G);
    const int k0 = gi * gsz, k1 = min(k0 + gsz, a.K);
    const float* Wr = a.W + (long)n * a.K;
    const float rs = a.row ? a.row[n] : 1.0f;
    const int rop = a.row ? a.row_op : 0;
    const float rfac = rop == 1 ? rs : (rop == 2 ? 1.0f / rs : 1.0f);
    const float en = a.e_row ? a.e_row[n] : 0.0f;
    const float xmin = a.xmin_in[item], xmax = a.xmax_in[item];
    const float su = lwc ? oq_sigmoidf(a.up[item]) : 1.0f;
    const float sl = lwc ? oq_sigmoidf(a.low[item]) : 1.0f;
    const bool ident = a.n_bits >= 16;   // fold only: dL/dW' = dL/dW_q, no scale path
    OqQParams q = {1.0f, 0.0f, 0.0f};
    if (!ident) q = oq_qparams(xmin, xmax, su, sl, lwc, a.n_bits, a.symmetric, a.recip_div);

    float acc[4] = {0.f, 0.f, 0.f, 0.f};  // R, #argmax, #argmin, row dot
    bool tie = false;
    for (int k = k0 + tr; k < k1; k += tsz) {
      const float w = Wr[k];
      const float w1 = has_col ? OQ_MUL(w, a.col[k]) : w;
      const float f = rop == 1 ? OQ_MUL(w1, rs) : (rop == 2 ? OQ_DIV(w1, rs) : w1);
      const float t = ident ? 0.0f : oq_code_unclamped(f, q);
      const bool inr = ident || ((t >= 0.0f) && (t <= q.qmax));
      const float c = oq_clampf(t, 0.0f, q.qmax);
      const float gk = load_g(a, (long)n * a.ld_g + k);
      if (!ident) acc[0] += gk * ((c - q.zp) - (inr ? f / q.scale : 0.0f));
      const bool ismax = !ident && (f == xmax), ismin = !ident && (f == xmin);
      acc[1] += ismax ? 1.0f : 0.0f;
      acc[2] += ismin ? 1.0f : 0.0f;
      tie |= ismax | ismin;
      const float d2 = inr ? gk : 0.0f;  // straight-through part of dL/dW2
      acc[3] += d2 * w1;
      if (colacc) {
        if (TEAM_WARP) atomicAdd(colacc + k, d2 * rfac * w);
        else colacc[k] += d2 * rfac * w;
      }
      if (shacc) {
        if (TEAM_WARP) atomicAdd(shacc + k, w * en);
        else shacc[k] += w * en;
      }
      if (a.dW) a.dW[(long)n * a.K + k] = d2;
    }
    if (k0 + gsz > a.K && tr == 0) {  // padded zeros count as ties of the extrema (their gradient is dropped)
      const float pad = (float)(k0 + gsz - a.K);
      if (xmax == 0.0f) acc[1] += pad;
      if (xmin == 0.0f) acc[2] += pad;
    }
    Team<TEAM_WARP>::reduce(acc, scratch, OpSum(), 0.0f);
    const float R = acc[0];
    // dL/dxmax', dL/dxmin' (after the LWC clip); symmetric variant through max(|.|,|.|) and the clamp
    float gmaxp = 0.0f, gminp = 0.0f;
    if (!ident) oq_scale_grad(R, xmin, xmax, su, sl, lwc, a.n_bits, a.symmetric, a.recip_div, &gmaxp, &gminp);
    if (tr == 0 && lwc) {
      if (a.d_up) a.d_up[item] = ident ? 0.0f : gmaxp * xmax * su * (1.0f - su);
      if (a.d_low) a.d_low[item] = ident ? 0.0f : gminp * xmin * sl * (1.0f - sl);
    }
    // gradient through amax/amin, split evenly over ties
    const float cmax = acc[1] > 0.f ? gmaxp * su / acc[1] : 0.0f;
    const float cmin = acc[2] > 0.f ? gminp * sl / acc[2] : 0.0f;
    float rowdot = 0.0f;
    if (tie && (a.dW || colacc || a.drow)) {
      for (int k = k0 + tr; k < k1; k += tsz) {
        const float w = Wr[k];
        const float w1 = has_col ? OQ_MUL(w, a.col[k]) : w;
        const float f = rop == 1 ? OQ_MUL(w1, rs) : (rop == 2 ? OQ_DIV(w1, rs) : w1);
        float corr = 0.0f;
        if (f == xmax) corr += cmax;
        if (f == xmin) corr += cmin;
        if (corr != 0.0f) {
          rowdot += corr * w1;
          if (colacc) {
            if (TEAM_WARP) atomicAdd(colacc + k, corr * rfac * w);
            else colacc[k] += corr * rfac * w;
          }
          if (a.dW) a.dW[(long)n * a.K + k] += corr;
        }
      }
    }
    if (a.drow) {
      float rd[1] = {rowdot};
      Team<TEAM_WARP>::reduce(rd, scratch, OpSum(), 0.0f);
      if (tr == 0) {
        const float v = acc[3] + rd[0];
        if (G == 1) a.drow[n] = v;
        else atomicAdd(a.drow + n, v);
      }
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < a.K; k += blockDim.x) {
    if (colacc && colacc[k] != 0.0f) atomicAdd(a.dcol + k, colacc[k]);
    if (shacc && shacc[k] != 0.0f) atomicAdd(a.dshift + k, shacc[k]);
  }
}

int pick_grid(long n_items, bool team_warp, int per_sm) {
  long teams_per_block = team_warp ? 8 : 1;
  long want = (n_items + teams_per_block - 1) / teams_per_block;
  long cap = (long)oq_num_sms() * per_sm;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

#include "weight_quant_fast.cuh"
#include "weight_quant_stream.cuh"

bool al(const void* p, int b) { return ((uintptr_t)p % b) == 0; }

}  // namespace

extern "C" int oq_weight_quant_fwd(const float* W, int N, int K, const float* col_scale, const float* row_scale,
                                   int row_op, const float* shift, float* matvec_out, int n_bits, int symmetric,
                                   int group_size, const float* up, const float* low, int recip_div, void* wq_bf16,
                                   long ld_bf16, float* wq_f32, uint8_t* codes, float* scale_out, float* zp_out,
                                   float* xmin_out, float* xmax_out, int8_t* codes_s8, float* s8_t, float* s8_beta,
                                   int n_seg, const int* seg_end_host, const float* const* seg_row_host,
                                   const int* seg_op_host, cudaStream_t stream) {
  OQ_REQUIRE((codes_s8 == nullptr) == (s8_t == nullptr) && (codes_s8 == nullptr) == (s8_beta == nullptr),
             "oq_weight_quant_fwd: codes_s8, s8_t and s8_beta go together");
  OQ_REQUIRE(!codes_s8 || (group_size == 0 && n_bits <= 8 && scale_out && (uintptr_t)codes_s8 % 4 == 0),
             "oq_weight_quant_fwd: the s8 side output needs per-channel quantisation, n_bits <= 8 and scale_out");
  OQ_REQUIRE(W && N > 0 && K > 0, "oq_weight_quant_fwd: bad shape N=%d K=%d", N, K);
  OQ_REQUIRE((n_bits >= 2 && n_bits <= 8) || n_bits >= 16, "oq_weight_quant_fwd: n_bits %d not in [2,8] or >= 16 (fold only)", n_bits);
  const bool ident = n_bits >= 16;
  OQ_REQUIRE(group_size >= 0, "oq_weight_quant_fwd: negative group_size");
  OQ_REQUIRE((up == nullptr) == (low == nullptr), "oq_weight_quant_fwd: up/low must both be set or both null");
  OQ_REQUIRE(group_size == 0 || K % group_size == 0 || symmetric,
             "oq_weight_quant_fwd: K %% group_size != 0 requires symmetric (quantizer.py:66-69)");
  OQ_REQUIRE(row_op >= 0 && row_op <= 2, "oq_weight_quant_fwd: bad row_op");
  WqArgs a = {};
  a.W = W; a.N = N; a.K = K; a.col = col_scale; a.row = row_scale; a.row_op = row_op; a.shift = shift; a.mv = matvec_out;
  a.n_bits = n_bits; a.symmetric = symmetric; a.group = group_size; a.recip_div = recip_div; a.up = up; a.low = low;
  a.q_bf = (bf16*)wq_bf16; a.ld_bf = ld_bf16 ? ld_bf16 : K; a.q_f32 = wq_f32; a.codes = codes;
  a.scale_out = scale_out; a.zp_out = zp_out; a.xmin_out = xmin_out; a.xmax_out = xmax_out;
  a.codes8 = codes_s8; a.s8_t = s8_t; a.s8_beta = s8_beta;
  OQ_REQUIRE(n_seg >= 0 && n_seg <= 3 && (n_seg == 0 || (seg_end_host && seg_row_host && seg_op_host && !row_scale)),
             "oq_weight_quant_fwd: bad row segment table (at most 3 segments; row_scale must be NULL with segments)");
  a.nseg = n_seg;
  for (int i = 0; i < n_seg; ++i) {
    a.seg_end[i] = seg_end_host[i]; a.seg_row[i] = seg_row_host[i]; a.seg_op[i] = seg_op_host[i];
    OQ_REQUIRE(a.seg_op[i] >= 0 && a.seg_op[i] <= 2 && a.seg_end[i] > (i ? a.seg_end[i - 1] : 0), "oq_weight_quant_fwd: bad segment %d", i);
  }
  OQ_REQUIRE(n_seg == 0 || (a.seg_end[n_seg - 1] == N && group_size == 0 && n_bits <= 8),
             "oq_weight_quant_fwd: segments must cover the N rows; per-channel quantisation only");
  const bool team_warp = group_size > 0;
  const long G = group_size ? (K + group_size - 1) / group_size : 1;
  const bool vec = (K % 4 == 0) && (group_size % 4 == 0) && (a.ld_bf % 4 == 0) && ((uintptr_t)W % 16 == 0) &&
                   ((uintptr_t)wq_bf16 % 8 == 0) && ((uintptr_t)wq_f32 % 16 == 0) && ((uintptr_t)codes % 4 == 0);
  if (shift && matvec_out && G > 1) {
    cudaError_t e = cudaMemsetAsync(matvec_out, 0, sizeof(float) * N, stream);
    if (e != cudaSuccess) return (int)e;
  }
  // fast register-resident paths (weight_quant_fast.cuh); the generic kernels below cover everything else
  const bool let = col_scale || row_scale || shift || n_seg;
  const bool fast_al = vec && al(col_scale, 16) && al(shift, 16);
  if (!ident && fast_al && !team_warp && K <= 16384) {
    const size_t smem = let ? sizeof(float) * 2 * K : 0;
    const int g = (int)std::min<long>(N, (long)oq_num_sms() * (K > 8192 ? 2 : 4));
    if (let ? launch_pc_fwd<true>(a, g, smem, stream) : launch_pc_fwd<false>(a, g, 0, stream))
      return oq_check_launch("oq_weight_quant_fwd");
  }
  if (!ident && fast_al && team_warp && K % group_size == 0 && (long)N * G < (1L << 31) - 64 && grp_lpg(group_size, false) >= 1) {
    const long slots = ((long)N * G * grp_lpg(group_size, false) + 31) / 32;
    const int g = (int)std::max<long>(1, std::min<long>((slots + 7) / 8, (long)oq_num_sms() * 8));
    if (let ? launch_grp<true, false>(a, g, 0, stream) : launch_grp<false, false>(a, g, 0, stream))
      return oq_check_launch("oq_weight_quant_fwd");
  }
  OQ_REQUIRE(!n_seg, "oq_weight_quant_fwd: row segments need the per-channel fast path (K %% 4 == 0, K <= 16384, aligned pointers)");
  OQ_REQUIRE(!codes_s8, "oq_weight_quant_fwd: the s8 side output needs the per-channel fast path (K %% 4 == 0, K <= 16384, "
                        "16-byte aligned W / col_scale / shift)");
  const int grid = pick_grid((long)N * G, team_warp, 8);
  if (team_warp) {
    if (vec) wq_fwd_kernel<true, true><<<grid, 256, 0, stream>>>(a);
    else wq_fwd_kernel<true, false><<<grid, 256, 0, stream>>>(a);
  } else {
    if (vec) wq_fwd_kernel<false, true><<<grid, 256, 0, stream>>>(a);
    else wq_fwd_kernel<false, false><<<grid, 256, 0, stream>>>(a);
  }
  return oq_check_launch("oq_weight_quant_fwd");
}

extern "C" int oq_weight_quant_bwd(const float* W, int N, int K, const float* col_scale, const float* row_scale,
                                   int row_op, int n_bits, int symmetric, int group_size, const float* up,
                                   const float* low, int recip_div, const float* xmin, const float* xmax,
                                   const float* scale_saved, const float* zp_saved, const void* grad_wq, int grad_is_f32, long ld_grad, const float* e_row,
                                   float* d_up, float* d_low, float* dcol_acc, float* drow, float* dshift_acc,
                                   float* dW_full, int n_seg, const int* seg_end_host, const float* const* seg_row_host,
                                   const int* seg_op_host, cudaStream_t stream) {
  OQ_REQUIRE(W && grad_wq && xmin && xmax && N > 0 && K > 0, "oq_weight_quant_bwd: null input or bad shape");
  OQ_REQUIRE((n_bits >= 2 && n_bits <= 8) || n_bits >= 16, "oq_weight_quant_bwd: n_bits %d not in [2,8] or >= 16 (fold only)", n_bits);
  const bool ident = n_bits >= 16;
  OQ_REQUIRE((up == nullptr) == (low == nullptr), "oq_weight_quant_bwd: up/low must both be set or both null");
  OQ_REQUIRE(!dshift_acc || e_row, "oq_weight_quant_bwd: dshift needs e_row");
  OQ_REQUIRE((scale_saved == nullptr) == (zp_saved == nullptr), "oq_weight_quant_bwd: scale/zp must both be given or both null");
  WqArgs a = {};
  a.W = W; a.N = N; a.K = K; a.col = col_scale; a.row = row_scale; a.row_op = row_op;
  a.n_bits = n_bits; a.symmetric = symmetric; a.group = group_size; a.recip_div = recip_div; a.up = up; a.low = low;
  a.scale_in = scale_saved; a.zp_in = zp_saved;
  a.g = grad_wq; a.g_is_f32 = grad_is_f32; a.ld_g = ld_grad ? ld_grad : K; a.xmin_in = xmin; a.xmax_in = xmax;
  a.e_row = e_row; a.d_up = d_up; a.d_low = d_low; a.dcol = dcol_acc; a.drow = drow; a.dshift = dshift_acc; a.dW = dW_full;
  OQ_REQUIRE(n_seg >= 0 && n_seg <= 3 && (n_seg == 0 || (seg_end_host && seg_row_host && seg_op_host && !row_scale)),
             "oq_weight_quant_bwd: bad row segment table (at most 3 segments; row_scale must be NULL with segments)");
  a.nseg = n_seg;
  for (int i = 0; i < n_seg; ++i) {
    a.seg_end[i] = seg_end_host[i]; a.seg_row[i] = seg_row_host[i]; a.seg_op[i] = seg_op_host[i];
    OQ_REQUIRE(a.seg_op[i] >= 0 && a.seg_op[i] <= 2 && a.seg_end[i] > (i ? a.seg_end[i - 1] : 0), "oq_weight_quant_bwd: bad segment %d", i);
  }
  OQ_REQUIRE(n_seg == 0 || (a.seg_end[n_seg - 1] == N && group_size == 0 && n_bits <= 8),
             "oq_weight_quant_bwd: segments must cover the N rows; per-channel quantisation only");
  const bool team_warp = group_size > 0;
  const long G = group_size ? (K + group_size - 1) / group_size : 1;
  const size_t smem = sizeof(float) * K * ((dcol_acc ? 1 : 0) + (dshift_acc ? 1 : 0));
  OQ_REQUIRE(smem <= 200 * 1024, "oq_weight_quant_bwd: K=%d too large for column accumulators", K);
  if (drow && G > 1) {
    cudaError_t e = cudaMemsetAsync(drow, 0, sizeof(float) * N, stream);
    if (e != cudaSuccess) return (int)e;
  }
  const bool let = col_scale || row_scale || e_row || dcol_acc || dshift_acc || n_seg;
  const bool fast_al = !ident && (K % 4 == 0) && (group_size % 4 == 0) && (a.ld_g % 4 == 0) && al(W, 16) && al(grad_wq, 16) &&
                       al(col_scale, 16) && al(dW_full, 16);
  if (fast_al && !team_warp && !let && !dW_full) {       // LWC only: R per row, one warp per row
    // one warp per row, two rows per block: small blocks keep every SM busy to the end (N = 4096 rows are 27.7 per SM)
    const int g = (int)std::min<long>((N + 1) / 2, (long)oq_num_sms() * 32);
    if (a.g_is_f32) wq_pc_bwd_lwc<4, true><<<g, 64, 0, stream>>>(a);
    else wq_pc_bwd_lwc<4, false><<<g, 64, 0, stream>>>(a);
    return oq_check_launch("oq_weight_quant_bwd");
  }
  if (fast_al && !team_warp && let && !dW_full && K % 8 == 0 && a.ld_g % 8 == 0) {   // bulk copies: 16-byte rows
    if (launch_pc_bwd_ws(a, stream)) return oq_check_launch("oq_weight_quant_bwd");
  }
  OQ_REQUIRE(!n_seg, "oq_weight_quant_bwd: row segments need the streaming per-channel path (K %% 8 == 0, K <= 8192, aligned pointers)");
  if (fast_al && !team_warp && K <= 16384) {
    const size_t sm2 = let ? sizeof(float) * 3 * K : 0;
    if (sm2 <= 200 * 1024) {
      const int g = (int)std::min<long>(N, (long)oq_num_sms() * (K > 4096 ? 2 : 3));
      if (let ? launch_pc_bwd<true>(a, g, sm2, stream) : launch_pc_bwd<false>(a, g, 0, stream))
        return oq_check_launch("oq_weight_quant_bwd");
    }
  }
  if (fast_al && team_warp && K % group_size == 0 && (long)N * G < (1L << 31) - 64 && grp_lpg(group_size, true) >= 1) {
    const size_t sm2 = let ? sizeof(float) * 2 * K : 0;
    const long slots = ((long)N * G * grp_lpg(group_size, true) + 31) / 32;
    const int g = (int)std::max<long>(1, std::min<long>((slots + 7) / 8, (long)oq_num_sms() * (let ? 3 : 8)));
    if (sm2 <= 200 * 1024 && (let ? launch_grp<true, true>(a, g, sm2, stream) : launch_grp<false, true>(a, g, 0, stream)))
      return oq_check_launch("oq_weight_quant_bwd");
  }
  // column accumulators are flushed once per block: fewer, longer-lived blocks when they are in use
  const int grid = pick_grid((long)N * G, team_warp, smem ? 2 : 8);
  if (team_warp) {
    if (smem > 40 * 1024) cudaFuncSetAttribute(wq_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    wq_bwd_kernel<true><<<grid, 256, smem, stream>>>(a);
  } else {
    if (smem > 40 * 1024) cudaFuncSetAttribute(wq_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    wq_bwd_kernel<false><<<grid, 256, smem, stream>>>(a);
  }
  return oq_check_launch("oq_weight_quant_bwd");
}
