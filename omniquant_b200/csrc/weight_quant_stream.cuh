// Streaming variants of the per-channel weight fake-quant backward (included by weight_quant.cu after
// weight_quant_fast.cuh).
//
// wq_pc_bwd_lwc : LWC only, no weight gradient wanted -- the backward reduces to R = dL/dscale per row, so one warp
//                 streams one row with U float4 loads in flight, no shared memory, no block barrier.
// wq_pc_bwd_pipe: LET.  Rows arrive in shared memory through 1-D bulk copies (cp.async.bulk + mbarrier, STAGES rows
//                 ahead per block), which takes the HBM latency off the critical path without holding a prefetched row in
//                 registers; the column accumulators of the LET gradients live in registers (a thread owns the same
//                 columns of every row its block visits) and are flushed with one atomicAdd per column per block.
#pragma once
#include "mbar.cuh"

template <int U>
__global__ void __launch_bounds__(256) wq_pc_bwd_lwc(WqArgs a) {
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = gridDim.x * (blockDim.x >> 5);
  const int K4 = a.K >> 2;
  const int lwc = a.up != nullptr;
  for (int n = wid; n < a.N; n += nw) {
    const float xmin = a.xmin_in[n], xmax = a.xmax_in[n];
    const float su = lwc ? oq_sigmoidf(a.up[n]) : 1.0f, sl = lwc ? oq_sigmoidf(a.low[n]) : 1.0f;
    OqQParams q;
    if (a.scale_in) {
      q.scale = a.scale_in[n];
      q.zp = a.zp_in[n];
      q.qmax = (float)((1 << a.n_bits) - 1);
    } else {
      q = oq_qparams(xmin, xmax, su, sl, lwc, a.n_bits, a.symmetric, a.recip_div);
    }
    const float inv = OQ_DIV(1.0f, q.scale);
    const float* wr = a.W + (long)n * a.K;
    float R = 0.f;
    for (int c0 = lane; c0 < K4; c0 += 32 * U) {
      float4 w[U], g[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int idx = c0 + 32 * u;
        if (idx < K4) {
          w[u] = ld4(wr + 4 * idx);
          if (a.g_is_f32) {
            g[u] = ld4(reinterpret_cast<const float*>(a.g) + (long)n * a.ld_g + 4 * idx);
          } else {
            float t[4];
            load_bf4(reinterpret_cast<const bf16*>(a.g) + (long)n * a.ld_g + 4 * idx, t);
            g[u] = make_float4(t[0], t[1], t[2], t[3]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (c0 + 32 * u < K4) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float qa = f4(w[u], j) * inv;
            const float t = rintf(qa) + q.zp;
            const float c = fminf(fmaxf(t, 0.0f), q.qmax);
            R += f4(g[u], j) * ((c - q.zp) - (c == t ? qa : 0.0f));
          }
        }
      }
    }
    R = warp_reduce(R, OpSum());
    if (lane == 0 && lwc) {
      float gmaxp, gminp;
      oq_scale_grad(R, xmin, xmax, su, sl, lwc, a.n_bits, a.symmetric, a.recip_div, &gmaxp, &gminp);
      if (a.d_up) a.d_up[n] = gmaxp * xmax * su * (1.0f - su);
      if (a.d_low) a.d_low[n] = gminp * xmin * sl * (1.0f - sl);
    }
  }
}

// dynamic smem: STAGES x { W row: K fp32 | dW_q row: K bf16 or fp32 }
// ROWDIV (the fold divides by the row scale: q and v): the forward's IEEE division is replaced by a reciprocal
// multiply, so the arg-max / arg-min elements are recognised within 2 ulp of the saved extrema; otherwise the fold is
// bit-identical to the forward's and the test is an equality.  Tie statistics go through (rare) shared-memory
// atomics, only R and the row dot product are reduced over the block.
template <int NV, int STAGES, bool ROWDIV>
__global__ void __launch_bounds__(256) wq_pc_bwd_pipe(WqArgs a) {
  extern __shared__ __align__(128) unsigned char psm[];
  __shared__ uint64_t full[STAGES];
  __shared__ float red[2][2][8];
  __shared__ float ties[3][4];   // #max, #min, sum w1@max, sum w1@min; triple-buffered over rows
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, K4 = a.K >> 2;
  const uint32_t wbytes = (uint32_t)a.K * 4u, gbytes = (uint32_t)a.K * (a.g_is_f32 ? 4u : 2u), sbytes = wbytes + gbytes;
  const size_t gsz = a.g_is_f32 ? 4 : 2;
  const bool has_col = a.col != nullptr, want_col = a.dcol != nullptr, want_sh = a.dshift != nullptr;
  const int lwc = a.up != nullptr;
  auto issue = [&](int row, int st) {
    mbar_expect_tx(&full[st], sbytes);
    bulk_g2s(psm + (size_t)st * sbytes, a.W + (long)row * a.K, wbytes, &full[st]);
    bulk_g2s(psm + (size_t)st * sbytes + wbytes, reinterpret_cast<const char*>(a.g) + (size_t)row * a.ld_g * gsz, gbytes,
             &full[st]);
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
    fence_proxy_async();
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      const long row = (long)blockIdx.x + (long)s * gridDim.x;
      if (row < a.N) issue((int)row, s);
    }
  }
  if (tid < 12) (&ties[0][0])[tid] = 0.f;
  float4 colr[NV], cacc[NV], sacc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = tid + i * 256;
    colr[i] = (has_col && idx < K4) ? ld4(a.col + 4 * idx) : make_float4(1.f, 1.f, 1.f, 1.f);
    cacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    sacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  int par = 0, it = 0, tb = 0;
  for (int n = blockIdx.x; n < a.N; n += gridDim.x, par ^= 1, ++it, tb = (tb == 2 ? 0 : tb + 1)) {
    const int st = it % STAGES;
    const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
    const float rs = (a.row && a.row_op != 0) ? a.row[n] : 1.0f;
    const float rmul = ROWDIV ? 1.0f / rs : rs;
    const float en = a.e_row ? a.e_row[n] : 0.0f;
    const float xmin = a.xmin_in[n], xmax = a.xmax_in[n];
    OqQParams q;
    if (a.scale_in) {
      q.scale = a.scale_in[n];
      q.zp = a.zp_in[n];
      q.qmax = (float)((1 << a.n_bits) - 1);
    } else {
      q = oq_qparams(xmin, xmax, lwc ? oq_sigmoidf(a.up[n]) : 1.0f, lwc ? oq_sigmoidf(a.low[n]) : 1.0f, lwc, a.n_bits,
                     a.symmetric, a.recip_div);
    }
    const float inv = OQ_DIV(1.0f, q.scale);
    const float tol_max = ROWDIV ? 2.4e-7f * fabsf(xmax) : 0.0f, tol_min = ROWDIV ? 2.4e-7f * fabsf(xmin) : 0.0f;
    mbar_wait(&full[st], ph);
    const float4* ws = reinterpret_cast<const float4*>(psm + (size_t)st * sbytes);
    const unsigned char* gs = psm + (size_t)st * sbytes + wbytes;
    float4 w[NV], g[NV];
    float R = 0.f, rowdot = 0.f;
    bool tie = false;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = tid + i * 256;
      if (idx < K4) {
        w[i] = ws[idx];
        if (a.g_is_f32) {
          g[i] = reinterpret_cast<const float4*>(gs)[idx];
        } else {
          const uint2 u = reinterpret_cast<const uint2*>(gs)[idx];
          unpack_bf2(u.x, g[i].x, g[i].y);
          unpack_bf2(u.y, g[i].z, g[i].w);
        }
        bool t4 = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float wv = f4(w[i], j);
          const float w1 = OQ_MUL(wv, f4(colr[i], j));
          const float f = OQ_MUL(w1, rmul);
          const float qa = f * inv;
          const float r = rintf(qa);
          const float t = r + q.zp;
          const float c = fminf(fmaxf(t, 0.0f), q.qmax);
          const bool inr = (c == t);
          const float gk = f4(g[i], j);
          R += gk * (inr ? (r - qa) : (c - q.zp));
          t4 |= ROWDIV ? (fabsf(f - xmax) <= tol_max) | (fabsf(f - xmin) <= tol_min) : (f == xmax) | (f == xmin);
          const float d2 = inr ? gk : 0.0f;
          rowdot += d2 * w1;
          f4(cacc[i], j) += (d2 * rmul) * wv;
          f4(sacc[i], j) += wv * en;
          f4(g[i], j) = d2;
        }
        if (t4) {   // rare: this float4 holds an arg-max / arg-min element of the row
          tie = true;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float w1 = OQ_MUL(f4(w[i], j), f4(colr[i], j));
            const float f = OQ_MUL(w1, rmul);
            if (ROWDIV ? fabsf(f - xmax) <= tol_max : f == xmax) { atomicAdd(&ties[tb][0], 1.0f); atomicAdd(&ties[tb][2], w1); }
            if (ROWDIV ? fabsf(f - xmin) <= tol_min : f == xmin) { atomicAdd(&ties[tb][1], 1.0f); atomicAdd(&ties[tb][3], w1); }
          }
        }
      }
    }
    R = warp_reduce(R, OpSum());
    rowdot = warp_reduce(rowdot, OpSum());
    if (lane == 0) {
      red[par][0][warp] = R;
      red[par][1][warp] = rowdot;
    }
    __syncthreads();
    // every thread has read this stage: refill it STAGES rows ahead; recycle the tie buffer two rows ahead
    if (tid == 0) {
      const long nn = (long)n + (long)STAGES * gridDim.x;
      if (nn < a.N) {
        fence_proxy_async();
        issue((int)nn, st);
      }
      const int tz = tb == 0 ? 2 : tb - 1;     // (tb + 2) % 3: last used one row ago, all its readers passed the previous barrier
      ties[tz][0] = 0.f; ties[tz][1] = 0.f; ties[tz][2] = 0.f; ties[tz][3] = 0.f;
    }
    if (tid == 0 || tie || a.dW) {
      float Rt = 0.f, rd = 0.f;
#pragma unroll
      for (int x = 0; x < 8; ++x) { Rt += red[par][0][x]; rd += red[par][1][x]; }
      const float nmax = ties[tb][0], nmin = ties[tb][1], smax = ties[tb][2], smin = ties[tb][3];
      const float su = lwc ? oq_sigmoidf(a.up[n]) : 1.0f;
      const float sl = lwc ? oq_sigmoidf(a.low[n]) : 1.0f;
      float gmaxp, gminp;
      oq_scale_grad(Rt, xmin, xmax, su, sl, lwc, a.n_bits, a.symmetric, a.recip_div, &gmaxp, &gminp);
      const float cmax = nmax > 0.f ? gmaxp * su / nmax : 0.0f;
      const float cmin = nmin > 0.f ? gminp * sl / nmin : 0.0f;
      if (tid == 0) {
        if (lwc && a.d_up) a.d_up[n] = gmaxp * xmax * su * (1.0f - su);
        if (lwc && a.d_low) a.d_low[n] = gminp * xmin * sl * (1.0f - sl);
        if (a.drow) a.drow[n] = rd + cmax * smax + cmin * smin;
      }
      if (tie || a.dW) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int idx = tid + i * 256;
          if (idx < K4) {
            float4 corr4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float wv = f4(w[i], j);
              const float f = OQ_MUL(OQ_MUL(wv, f4(colr[i], j)), rmul);
              float corr = 0.0f;
              if (ROWDIV ? fabsf(f - xmax) <= tol_max : f == xmax) corr += cmax;
              if (ROWDIV ? fabsf(f - xmin) <= tol_min : f == xmin) corr += cmin;
              f4(corr4, j) = corr;
              f4(cacc[i], j) += corr * rmul * wv;
            }
            if (a.dW)
              *reinterpret_cast<float4*>(a.dW + (long)n * a.K + 4 * idx) =
                  make_float4(g[i].x + corr4.x, g[i].y + corr4.y, g[i].z + corr4.z, g[i].w + corr4.w);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = tid + i * 256;
    if (idx < K4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (want_col) atomicAdd(a.dcol + 4 * idx + j, f4(cacc[i], j));
        if (want_sh) atomicAdd(a.dshift + 4 * idx + j, f4(sacc[i], j));
      }
    }
  }
}

template <bool ROWDIV>
static bool launch_pc_bwd_pipe2(const WqArgs& a, cudaStream_t s) {
  constexpr int STAGES = 3;
  const int nv = (a.K / 4 + 255) / 256;
  const size_t smem = (size_t)STAGES * a.K * (a.g_is_f32 ? 8 : 6);
  if (nv > 8 || smem > 200 * 1024) return false;
  const int bps = (int)std::min<size_t>(nv <= 2 ? 3 : (nv <= 4 ? 2 : 1), (220 * 1024) / (smem + 1024));   // registers per thread bound it too
  const int grid = (int)std::min<long>(a.N, (long)oq_num_sms() * bps);
#define OQ_PIPE(NVV)                                                                                                  \
  {                                                                                                                   \
    cudaFuncSetAttribute(wq_pc_bwd_pipe<NVV, STAGES, ROWDIV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    wq_pc_bwd_pipe<NVV, STAGES, ROWDIV><<<grid, 256, smem, s>>>(a);                                                   \
    return true;                                                                                                      \
  }
  if (nv <= 1) OQ_PIPE(1)
  if (nv <= 2) OQ_PIPE(2)
  if (nv <= 4) OQ_PIPE(4)
  if (nv <= 8) OQ_PIPE(8)
#undef OQ_PIPE
  return false;
}
static bool launch_pc_bwd_pipe(const WqArgs& a, cudaStream_t s) {
  return (a.row && a.row_op == 2) ? launch_pc_bwd_pipe2<true>(a, s) : launch_pc_bwd_pipe2<false>(a, s);
}
