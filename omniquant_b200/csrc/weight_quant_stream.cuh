// Streaming variants of the per-channel weight fake-quant backward (included by weight_quant.cu after
// weight_quant_fast.cuh).
//
// wq_pc_bwd_lwc : LWC only, no weight gradient wanted -- the backward reduces to R = dL/dscale per row, so one warp
//                 streams one row with U float4 loads in flight, no shared memory, no block barrier.
// wq_pc_bwd_ws  : LET.  Warp-specialised: rows arrive in shared memory through 1-D bulk copies (cp.async.bulk +
//                 mbarrier, STAGES rows ahead per block) issued by a control warp; the column accumulators of the LET
//                 gradients live in registers (a thread owns the same columns of every row its block visits) and are
//                 flushed with one atomicAdd per column per block.  ROWDIV (the fold divides by the row scale: q and
//                 v): the forward's IEEE division is replaced by a reciprocal multiply, so the arg-max / arg-min
//                 elements are recognised within 2 ulp of the saved extrema; otherwise the fold is bit-identical to
//                 the forward's and the test is an equality.
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

// ------------------------------------------------------------------ warp-specialised LET backward
// 8 consumer warps + 1 control warp, no block barrier in the row loop.  The control warp owns everything that is per
// row: it prefetches the row's scalars, derives the header the consumers need (1/scale, fold factor, extrema), issues
// the bulk copies of the row into a shared-memory stage, and -- once all consumer warps have released the stage --
// reduces their partial sums, evaluates the clip-path gradient, applies it to the (few) arg-max / arg-min elements the
// consumers reported in a per-stage list and writes d_up / d_low / drow.  Consumers only wait on the stage's `full`
// barrier, so a slow row finalisation or a late row never stalls the other warps.
//   smem: STAGES x { W row: K fp32 | dW_q row: K bf16 or fp32 }
constexpr int WS_TIE_CAP = 32;
struct WsRowHdr {
  float rmul, en, xmin, xmax, zp, inv, qmax, su, sl, tolmax, tolmin, pad;
};

// ROWMODE: 0 the row fold multiplies (or is absent), 1 it divides, 2 per-row segment table (q | k | v in one launch)
template <int NV, int STAGES, int ROWMODE>
__global__ void __launch_bounds__(288, NV <= 4 ? 2 : 1) wq_pc_bwd_ws(WqArgs a) {
  constexpr bool ROWDIV = ROWMODE != 0;      // consumers: recognise the extrema within the row's tolerance (0 when it multiplies)
  extern __shared__ __align__(128) unsigned char psm[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  __shared__ WsRowHdr hdr[STAGES];
  __shared__ float red[STAGES][2][8];
  __shared__ int tie_cnt[STAGES];
  __shared__ int tie_k[STAGES][WS_TIE_CAP];
  __shared__ float tie_wv[STAGES][WS_TIE_CAP], tie_w1[STAGES][WS_TIE_CAP];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, K4 = a.K >> 2;
  const uint32_t wbytes = (uint32_t)a.K * 4u, gbytes = (uint32_t)a.K * (a.g_is_f32 ? 4u : 2u), sbytes = wbytes + gbytes;
  const size_t gsz = a.g_is_f32 ? 4 : 2;
  const int lwc = a.up != nullptr;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 8);
      tie_cnt[s] = 0;
    }
    fence_mbar_init();
    fence_proxy_async();
  }
  __syncthreads();

  if (warp == 8) {
    // ---------------------------------------------------------------- control warp
    struct Raw { float rs, en, xmin, xmax, scale, zp, up, low; int div; };
    auto load_raw = [&](long row) {
      Raw r;
      if (ROWMODE == 2) {
        int op;
        wq_row_fold(a, row, r.rs, op);
        r.div = op == 2;
      } else {
        r.rs = (a.row && a.row_op != 0) ? a.row[row] : 1.0f;
        r.div = ROWMODE == 1;
      }
      r.en = a.e_row ? a.e_row[row] : 0.0f;
      r.xmin = a.xmin_in[row];
      r.xmax = a.xmax_in[row];
      r.scale = a.scale_in ? a.scale_in[row] : 0.0f;
      r.zp = a.zp_in ? a.zp_in[row] : 0.0f;
      r.up = lwc ? a.up[row] : 0.0f;
      r.low = lwc ? a.low[row] : 0.0f;
      return r;
    };
    auto fill = [&](long row, int st, const Raw& r) {   // lane 0: header, then the row itself
      WsRowHdr h;
      h.su = lwc ? oq_sigmoidf(r.up) : 1.0f;
      h.sl = lwc ? oq_sigmoidf(r.low) : 1.0f;
      OqQParams q;
      if (a.scale_in) {
        q.scale = r.scale;
        q.zp = r.zp;
        q.qmax = (float)((1 << a.n_bits) - 1);
      } else {
        q = oq_qparams(r.xmin, r.xmax, h.su, h.sl, lwc, a.n_bits, a.symmetric, a.recip_div);
      }
      h.rmul = r.div ? 1.0f / r.rs : r.rs;
      h.en = r.en; h.xmin = r.xmin; h.xmax = r.xmax; h.zp = q.zp; h.qmax = q.qmax;
      h.inv = OQ_DIV(1.0f, q.scale);
      h.tolmax = r.div ? 2.4e-7f * fabsf(r.xmax) : 0.0f;
      h.tolmin = r.div ? 2.4e-7f * fabsf(r.xmin) : 0.0f;
      h.pad = 0.f;
      hdr[st] = h;
      fence_proxy_async();
      mbar_expect_tx(&full[st], sbytes);     // release: the header is visible to whoever observes the phase
      bulk_g2s(psm + (size_t)st * sbytes, a.W + row * a.K, wbytes, &full[st]);
      bulk_g2s(psm + (size_t)st * sbytes + wbytes, reinterpret_cast<const char*>(a.g) + (size_t)row * a.ld_g * gsz, gbytes,
               &full[st]);
    };
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < STAGES; ++s) {
        const long row = (long)blockIdx.x + (long)s * gridDim.x;
        if (row < a.N) fill(row, s, load_raw(row));
      }
    }
    int it = 0;
    for (long n = blockIdx.x; n < a.N; n += gridDim.x, ++it) {
      const int st = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      const long nn = n + (long)STAGES * gridDim.x;
      Raw nxt = {};
      if (lane == 0 && nn < a.N) nxt = load_raw(nn);      // in flight while the consumers finish row n
      mbar_wait(&empty[st], ph);
      const WsRowHdr h = hdr[st];
      float Rt = lane < 8 ? red[st][0][lane] : 0.f, rd = lane < 8 ? red[st][1][lane] : 0.f;
      Rt = warp_reduce(Rt, OpSum());
      rd = warp_reduce(rd, OpSum());
      const int cnt = tie_cnt[st];
      float gmaxp, gminp;
      oq_scale_grad(Rt, h.xmin, h.xmax, h.su, h.sl, lwc, a.n_bits, a.symmetric, a.recip_div, &gmaxp, &gminp);
      float nmax = 0.f, nmin = 0.f, smax = 0.f, smin = 0.f;
      const float4* ws = reinterpret_cast<const float4*>(psm + (size_t)st * sbytes);
      if (cnt <= WS_TIE_CAP) {
        const bool have = lane < cnt;
        const int kf = have ? tie_k[st][lane] : 0;
        const float w1 = have ? tie_w1[st][lane] : 0.f;
        nmax = (float)__popc(__ballot_sync(0xffffffffu, have && (kf & 1)));
        nmin = (float)__popc(__ballot_sync(0xffffffffu, have && (kf & 2)));
        smax = warp_reduce((kf & 1) ? w1 : 0.f, OpSum());
        smin = warp_reduce((kf & 2) ? w1 : 0.f, OpSum());
        const float cmax = nmax > 0.f ? gmaxp * h.su / nmax : 0.0f, cmin = nmin > 0.f ? gminp * h.sl / nmin : 0.0f;
        if (have && a.dcol) {
          const float corr = ((kf & 1) ? cmax : 0.f) + ((kf & 2) ? cmin : 0.f);
          atomicAdd(a.dcol + (kf >> 2), corr * h.rmul * tie_wv[st][lane]);
        }
      } else {   // many equal extrema (e.g. already-quantised weights): rescan the row, which is still in the stage
        for (int pass = 0; pass < 2; ++pass) {
          const float cmax = nmax > 0.f ? gmaxp * h.su / nmax : 0.0f, cmin = nmin > 0.f ? gminp * h.sl / nmin : 0.0f;
          for (int idx = lane; idx < K4; idx += 32) {
            const float4 w4 = ws[idx];
            const float4 c4 = a.col ? ld4(a.col + 4 * idx) : make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float wv = f4(const_cast<float4&>(w4), j);
              const float w1 = OQ_MUL(wv, f4(const_cast<float4&>(c4), j));
              const float f = OQ_MUL(w1, h.rmul);
              const bool ismax = ROWDIV ? fabsf(f - h.xmax) <= h.tolmax : f == h.xmax;
              const bool ismin = ROWDIV ? fabsf(f - h.xmin) <= h.tolmin : f == h.xmin;
              if (pass == 0) {
                if (ismax) { nmax += 1.0f; smax += w1; }
                if (ismin) { nmin += 1.0f; smin += w1; }
              } else if ((ismax || ismin) && a.dcol) {
                atomicAdd(a.dcol + 4 * idx + j, ((ismax ? cmax : 0.f) + (ismin ? cmin : 0.f)) * h.rmul * wv);
              }
            }
          }
          if (pass == 0) {
            nmax = warp_reduce(nmax, OpSum()); nmin = warp_reduce(nmin, OpSum());
            smax = warp_reduce(smax, OpSum()); smin = warp_reduce(smin, OpSum());
          }
        }
      }
      if (lane == 0) {
        const float cmax = nmax > 0.f ? gmaxp * h.su / nmax : 0.0f, cmin = nmin > 0.f ? gminp * h.sl / nmin : 0.0f;
        if (lwc && a.d_up) a.d_up[n] = gmaxp * h.xmax * h.su * (1.0f - h.su);
        if (lwc && a.d_low) a.d_low[n] = gminp * h.xmin * h.sl * (1.0f - h.sl);
        if (a.drow) a.drow[n] = rd + cmax * smax + cmin * smin;
        tie_cnt[st] = 0;
      }
      __syncwarp();
      if (lane == 0 && nn < a.N) fill(nn, st, nxt);
    }
    return;
  }

  // ------------------------------------------------------------------ consumer warps
  const bool want_col = a.dcol != nullptr, want_sh = a.dshift != nullptr;
  float4 colr[NV], cacc[NV], sacc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = tid + i * 256;
    colr[i] = (a.col && idx < K4) ? ld4(a.col + 4 * idx) : make_float4(1.f, 1.f, 1.f, 1.f);
    cacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    sacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int it = 0;
  for (long n = blockIdx.x; n < a.N; n += gridDim.x, ++it) {
    const int st = it % STAGES;
    const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
    mbar_wait(&full[st], ph);
    const WsRowHdr h = hdr[st];
    const float4* ws = reinterpret_cast<const float4*>(psm + (size_t)st * sbytes);
    const unsigned char* gs = psm + (size_t)st * sbytes + wbytes;
    float R = 0.f, rowdot = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = tid + i * 256;
      if (idx < K4) {
        const float4 w4 = ws[idx];
        float4 g4;
        if (a.g_is_f32) {
          g4 = reinterpret_cast<const float4*>(gs)[idx];
        } else {
          const uint2 u = reinterpret_cast<const uint2*>(gs)[idx];
          unpack_bf2(u.x, g4.x, g4.y);
          unpack_bf2(u.y, g4.z, g4.w);
        }
        bool t4 = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float wv = f4(const_cast<float4&>(w4), j);
          const float w1 = OQ_MUL(wv, f4(colr[i], j));
          const float f = OQ_MUL(w1, h.rmul);
          const float qa = f * h.inv;
          const float r = rintf(qa);
          const float t = r + h.zp;
          const float c = fminf(fmaxf(t, 0.0f), h.qmax);
          const bool inr = (c == t);
          const float gk = f4(g4, j);
          R += gk * (inr ? (r - qa) : (c - h.zp));
          t4 |= ROWDIV ? (fabsf(f - h.xmax) <= h.tolmax) | (fabsf(f - h.xmin) <= h.tolmin) : (f == h.xmax) | (f == h.xmin);
          const float d2 = inr ? gk : 0.0f;
          rowdot += d2 * w1;
          f4(cacc[i], j) += (d2 * h.rmul) * wv;
          f4(sacc[i], j) += wv * h.en;
        }
        if (t4) {   // rare: report the arg-max / arg-min elements of the row to the control warp
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float wv = f4(const_cast<float4&>(w4), j);
            const float w1 = OQ_MUL(wv, f4(colr[i], j));
            const float f = OQ_MUL(w1, h.rmul);
            const bool ismax = ROWDIV ? fabsf(f - h.xmax) <= h.tolmax : f == h.xmax;
            const bool ismin = ROWDIV ? fabsf(f - h.xmin) <= h.tolmin : f == h.xmin;
            if (ismax || ismin) {
              const int slot = atomicAdd(&tie_cnt[st], 1);
              if (slot < WS_TIE_CAP) {
                tie_k[st][slot] = ((4 * idx + j) << 2) | (ismax ? 1 : 0) | (ismin ? 2 : 0);
                tie_wv[st][slot] = wv;
                tie_w1[st][slot] = w1;
              }
            }
          }
        }
      }
    }
    R = warp_reduce(R, OpSum());
    rowdot = warp_reduce(rowdot, OpSum());
    if (lane == 0) {
      red[st][0][warp] = R;
      red[st][1][warp] = rowdot;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);    // release: partial sums, tie list and the stage itself
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

template <int ROWDIV>
static bool launch_pc_bwd_ws2(const WqArgs& a, cudaStream_t s) {
  constexpr int STAGES = 3;
  const int nv = (a.K / 4 + 255) / 256;
  const size_t smem = (size_t)STAGES * a.K * (a.g_is_f32 ? 8 : 6);
  if (nv > 8 || smem > 200 * 1024 || a.dW) return false;
  const int bps = (int)std::min<size_t>(nv <= 4 ? 2 : 1, (220 * 1024) / (smem + 2048));
  const int grid = (int)std::min<long>(a.N, (long)oq_num_sms() * bps);
#define OQ_WS(NVV)                                                                                                  \
  {                                                                                                                 \
    cudaFuncSetAttribute(wq_pc_bwd_ws<NVV, STAGES, ROWDIV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    wq_pc_bwd_ws<NVV, STAGES, ROWDIV><<<grid, 288, smem, s>>>(a);                                                   \
    return true;                                                                                                    \
  }
  if (nv <= 1) OQ_WS(1)
  if (nv <= 2) OQ_WS(2)
  if (nv <= 4) OQ_WS(4)
  if (nv <= 8) OQ_WS(8)
#undef OQ_WS
  return false;
}
static bool launch_pc_bwd_ws(const WqArgs& a, cudaStream_t s) {
  if (a.nseg) return launch_pc_bwd_ws2<2>(a, s);
  return (a.row && a.row_op == 2) ? launch_pc_bwd_ws2<1>(a, s) : launch_pc_bwd_ws2<0>(a, s);
}
