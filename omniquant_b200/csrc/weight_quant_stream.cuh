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

template <int U, bool G32>
__global__ void __launch_bounds__(64) wq_pc_bwd_lwc(WqArgs a) {
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = gridDim.x * (blockDim.x >> 5);
  const int K4 = a.K >> 2;
  const int lwc = a.up != nullptr;
  const float magic = 12582912.0f;
  using GV = typename std::conditional<G32, float4, uint2>::type;     // four gradients as loaded
  for (int n = wid; n < a.N; n += nw) {
    const float* wr = a.W + (long)n * a.K;
    const GV* gr = reinterpret_cast<const GV*>(reinterpret_cast<const char*>(a.g) + (size_t)n * a.ld_g * (G32 ? 4 : 2));
    // the row's first chunk is in flight while the row scalars are evaluated; chunk c + 1 is in flight while chunk c is
    // consumed (ncu on the unpipelined loop: 12 warp-cycles of long-scoreboard stall per issued instruction)
    float4 w[U], wn[U];
    GV g[U], gn[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = lane + 32 * u;
      if (idx < K4) {
        wn[u] = ld4(wr + 4 * idx);
        gn[u] = __ldg(gr + idx);
      }
    }
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
    // in range (rint(qa) + zp in [0, qmax]) <=> qlo <= qa <= qhi, see WsRowHdr
    const float lo = -q.zp, hi = q.qmax - q.zp;
    const float qlo = fmodf(lo, 2.0f) == 0.0f ? lo - 0.5f : nextafterf(lo - 0.5f, INFINITY);
    const float qhi = fmodf(hi, 2.0f) == 0.0f ? hi + 0.5f : nextafterf(hi + 0.5f, -INFINITY);
    float R = 0.f;
    for (int c0 = lane; c0 < K4; c0 += 32 * U) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        w[u] = wn[u];
        g[u] = gn[u];
      }
      const int c1 = c0 + 32 * U;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int idx = c1 + 32 * u;
        if (idx < K4) {
          wn[u] = ld4(wr + 4 * idx);
          gn[u] = __ldg(gr + idx);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (c0 + 32 * u < K4) {
          float g4[4];
          if constexpr (G32) {
            g4[0] = g[u].x; g4[1] = g[u].y; g4[2] = g[u].z; g4[3] = g[u].w;
          } else {
            unpack_bf2(g[u].x, g4[0], g4[1]);
            unpack_bf2(g[u].y, g4[2], g4[3]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float qa = f4(w[u], j) * inv;
            const float qc = fminf(fmaxf(qa, qlo), qhi);
            const float rc = __fsub_rn(__fadd_rn(qc, magic), magic);     // clamp(rint(qa), -zp, qmax - zp)
            R = fmaf(g4[j], rc - (qc == qa ? qa : 0.0f), R);
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
constexpr int WS_FIN_TIES = 4;     // extremal elements per row that the batched finalisation keeps (more: the row is finalised at once)
struct WsRowHdr {
  float rmul, en, xmin, xmax, zp, inv, qmax, su, sl, tolmax, tolmin, pad;
  // rint(q) + zp lies in [0, qmax]  <=>  qlo <= q <= qhi: the half-integers -zp - 0.5 / qmax - zp + 0.5, moved one ulp
  // inwards when the tie rounds (to even) out of the range
  float qlo, qhi, pad2, pad3;
};

// ROWMODE: 0 the row fold multiplies (or is absent), 1 it divides, 2 per-row segment table (q | k | v in one launch)
template <int NV, int STAGES, int ROWMODE, bool G32, bool EXACT>
__global__ void __launch_bounds__(288, NV <= 4 ? 2 : 1) wq_pc_bwd_ws(WqArgs a) {
  constexpr bool ROWDIV = ROWMODE != 0;      // consumers: recognise the extrema within the row's tolerance (0 when it multiplies)
  extern __shared__ __align__(128) unsigned char psm[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  __shared__ WsRowHdr hdr[STAGES];
  __shared__ float red[STAGES][2][8];
  __shared__ int tie_cnt[STAGES];
  __shared__ int tie_k[STAGES][WS_TIE_CAP];
  __shared__ float tie_wv[STAGES][WS_TIE_CAP], tie_w1[STAGES][WS_TIE_CAP];
  // control warp: prepared headers and parked row results of two 32-row chunks
  __shared__ WsRowHdr htab[2][32];
  __shared__ float fin_R[2][32], fin_rd[2][32];
  __shared__ int fin_cnt[2][32];
  __shared__ int fin_k[2][32][WS_FIN_TIES];
  __shared__ float fin_wv[2][32][WS_FIN_TIES], fin_w1[2][32][WS_FIN_TIES];
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
    // Everything that is per row and scalar (two sigmoids, 1/scale, the range bounds before the row; the clip-path
    // gradient, d_up / d_low / drow and the extrema corrections after it) is evaluated for 32 rows AT ONCE, one row per
    // lane: done row by row on one lane it cost ~2 us per row and bounded the whole kernel (ncu: the consumers spent 36 %
    // of their samples spinning on the stage barrier, more stages did not help, one block per SM halved the speed).  The
    // serial per-row path is now: wait for the stage, fold the eight warp partials, park them in the row's slot, copy
    // the prepared header and re-arm the bulk copies.
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
    auto header = [&](const Raw& r) {
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
      h.pad = h.pad2 = h.pad3 = 0.f;
      const float lo = -q.zp, hi = q.qmax - q.zp;          // integers (|zp| <= 1e4): the sums below are exact
      const bool lo_even = fmodf(lo, 2.0f) == 0.0f, hi_even = fmodf(hi, 2.0f) == 0.0f;
      h.qlo = lo_even ? lo - 0.5f : nextafterf(lo - 0.5f, INFINITY);
      h.qhi = hi_even ? hi + 0.5f : nextafterf(hi + 0.5f, -INFINITY);
      return h;
    };
    const int grid = gridDim.x;
    const int nrows = (a.N - (int)blockIdx.x + grid - 1) / grid;      // rows of this block: blockIdx.x + i * grid
    auto row_of = [&](int i) { return (long)blockIdx.x + (long)i * grid; };
    auto headers = [&](int chunk) {      // all lanes: headers of rows [32 chunk, 32 chunk + 32)
      const int i = 32 * chunk + lane;
      if (i < nrows) htab[chunk & 1][lane] = header(load_raw(row_of(i)));
      __syncwarp();
    };
    auto fill = [&](int i, int st) {   // lane 0: publish row i's header, then fetch the row itself
      const long row = row_of(i);
      hdr[st] = htab[(i >> 5) & 1][i & 31];
      fence_proxy_async();
      mbar_expect_tx(&full[st], sbytes);     // release: the header is visible to whoever observes the phase
      bulk_g2s(psm + (size_t)st * sbytes, a.W + row * a.K, wbytes, &full[st]);
      bulk_g2s(psm + (size_t)st * sbytes + wbytes, reinterpret_cast<const char*>(a.g) + (size_t)row * a.ld_g * gsz, gbytes,
               &full[st]);
    };
    // one row, one lane: clip-path gradient, corrections of the (few) extremal elements, outputs
    auto finalize = [&](int i, const WsRowHdr& h, float Rt, float rd, int cnt, const int* tk, const float* twv, const float* tw1) {
      float gmaxp, gminp;
      oq_scale_grad(Rt, h.xmin, h.xmax, h.su, h.sl, lwc, a.n_bits, a.symmetric, a.recip_div, &gmaxp, &gminp);
      float nmax = 0.f, nmin = 0.f, smax = 0.f, smin = 0.f;
      for (int t = 0; t < cnt; ++t) {
        if (tk[t] & 1) { nmax += 1.0f; smax += tw1[t]; }
        if (tk[t] & 2) { nmin += 1.0f; smin += tw1[t]; }
      }
      const float cmax = nmax > 0.f ? gmaxp * h.su / nmax : 0.0f, cmin = nmin > 0.f ? gminp * h.sl / nmin : 0.0f;
      if (a.dcol)
        for (int t = 0; t < cnt; ++t)
          atomicAdd(a.dcol + (tk[t] >> 2), (((tk[t] & 1) ? cmax : 0.f) + ((tk[t] & 2) ? cmin : 0.f)) * h.rmul * twv[t]);
      const long n = row_of(i);
      if (lwc && a.d_up) a.d_up[n] = gmaxp * h.xmax * h.su * (1.0f - h.su);
      if (lwc && a.d_low) a.d_low[n] = gminp * h.xmin * h.sl * (1.0f - h.sl);
      if (a.drow) a.drow[n] = rd + cmax * smax + cmin * smin;
    };
    headers(0);
    if (lane == 0)
      for (int i = 0; i < STAGES && i < nrows; ++i) fill(i, i);
    int st = 0;
    uint32_t ph = 0;
    for (int i = 0; i < nrows; ++i) {
      const int j = i + STAGES;                 // the row that takes over this stage
      const int slot = i & 31, fb = (i >> 5) & 1;
      mbar_wait(&empty[st], ph);
      float Rt = lane < 8 ? red[st][0][lane] : 0.f, rd = lane < 8 ? red[st][1][lane] : 0.f;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        Rt += __shfl_xor_sync(0xffffffffu, Rt, o);
        rd += __shfl_xor_sync(0xffffffffu, rd, o);
      }
      const int cnt = tie_cnt[st];
      if (cnt <= WS_FIN_TIES) {               // the common case: park the row's partials, hand the stage back
        if (lane < cnt) {
          fin_k[fb][slot][lane] = tie_k[st][lane];
          fin_wv[fb][slot][lane] = tie_wv[st][lane];
          fin_w1[fb][slot][lane] = tie_w1[st][lane];
        }
        if (lane == 0) {
          fin_R[fb][slot] = Rt;
          fin_rd[fb][slot] = rd;
          fin_cnt[fb][slot] = cnt;
        }
        __syncwarp();
        if (lane == 0) {
          tie_cnt[st] = 0;
          if (j < nrows) fill(j, st);
        }
      } else {                                // many equal extrema: finalise this row now, the stage still holds it
        const WsRowHdr h = hdr[st];
        float gmaxp, gminp;
        oq_scale_grad(Rt, h.xmin, h.xmax, h.su, h.sl, lwc, a.n_bits, a.symmetric, a.recip_div, &gmaxp, &gminp);
        float nmax = 0.f, nmin = 0.f, smax = 0.f, smin = 0.f;
        const float4* ws = reinterpret_cast<const float4*>(psm + (size_t)st * sbytes);
        for (int pass = 0; pass < 2; ++pass) {
          const float cmax = nmax > 0.f ? gmaxp * h.su / nmax : 0.0f, cmin = nmin > 0.f ? gminp * h.sl / nmin : 0.0f;
          for (int idx = lane; idx < K4; idx += 32) {
            const float4 w4 = ws[idx];
            const float4 c4 = a.col ? ld4(a.col + 4 * idx) : make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float wv = f4(const_cast<float4&>(w4), e);
              const float w1s = OQ_MUL(wv, f4(const_cast<float4&>(c4), e));
              const float f = OQ_MUL(w1s, h.rmul);
              const bool ismax = ROWDIV ? fabsf(f - h.xmax) <= h.tolmax : f == h.xmax;
              const bool ismin = ROWDIV ? fabsf(f - h.xmin) <= h.tolmin : f == h.xmin;
              if (pass == 0) {
                if (ismax) { nmax += 1.0f; smax += w1s; }
                if (ismin) { nmin += 1.0f; smin += w1s; }
              } else if ((ismax || ismin) && a.dcol) {
                atomicAdd(a.dcol + 4 * idx + e, ((ismax ? cmax : 0.f) + (ismin ? cmin : 0.f)) * h.rmul * wv);
              }
            }
          }
          if (pass == 0) {
            nmax = warp_reduce(nmax, OpSum()); nmin = warp_reduce(nmin, OpSum());
            smax = warp_reduce(smax, OpSum()); smin = warp_reduce(smin, OpSum());
          }
        }
        __syncwarp();
        if (lane == 0) {
          const float cmax = nmax > 0.f ? gmaxp * h.su / nmax : 0.0f, cmin = nmin > 0.f ? gminp * h.sl / nmin : 0.0f;
          const long n = row_of(i);
          if (lwc && a.d_up) a.d_up[n] = gmaxp * h.xmax * h.su * (1.0f - h.su);
          if (lwc && a.d_low) a.d_low[n] = gminp * h.xmin * h.sl * (1.0f - h.sl);
          if (a.drow) a.drow[n] = rd + cmax * smax + cmin * smin;
          fin_cnt[fb][slot] = -1;             // done
          tie_cnt[st] = 0;
          if (j < nrows) fill(j, st);
        }
      }
      if (++st == STAGES) { st = 0; ph ^= 1u; }
      // headers of the next chunk: its first fill is STAGES rows before the chunk starts, and the table it replaces
      // belongs to the chunk that was finalised at the end of the previous 32 rows
      if (slot == 8 && 32 * ((i >> 5) + 1) < nrows) headers((i >> 5) + 1);
      if (slot == 31 || i == nrows - 1) {     // a chunk is complete: finalise its rows, one per lane
        __syncwarp();
        const int cntl = lane <= slot ? fin_cnt[fb][lane] : -1;
        if (cntl >= 0) finalize(i - slot + lane, htab[fb][lane], fin_R[fb][lane], fin_rd[fb][lane], cntl, fin_k[fb][lane], fin_wv[fb][lane], fin_w1[fb][lane]);
        __syncwarp();
      }
    }
    return;
  }

  // ------------------------------------------------------------------ consumer warps
  // Instruction diet (ncu: 34.5 thread-instructions per weight before, ~11 of them integer / control): gradient type and
  // "K fills the threads exactly" are template parameters, so bounds tests apply to a thread's LAST float4 only; the
  // in-range test and the clamp are one operation -- qc = clamp(q, qlo, qhi) with the half-integer bounds of the header,
  // in range <=> qc == q, and rint(qc) (magic-number add) is the clamped code minus the zero point; the arg-max / arg-min
  // probe is one 3-input max / min pair per float4; the two row sums share one butterfly (5 shuffles instead of 10).
  const bool want_col = a.dcol != nullptr, want_sh = a.dshift != nullptr;
  const bool last_ok = tid + (NV - 1) * 256 < K4;
  const int nfull = (K4 - tid + 255) / 256;
#define OQ_WS_OK(i) (EXACT ? ((i) < NV - 1 || last_ok) : ((i) < nfull))
  float4 colr[NV], cacc[NV], sacc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    colr[i] = (a.col && OQ_WS_OK(i)) ? ld4(a.col + 4 * (tid + i * 256)) : make_float4(1.f, 1.f, 1.f, 1.f);
    cacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    sacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float magic = 12582912.0f;
  int it = 0, st = 0;
  uint32_t ph = 0;
  for (long n = blockIdx.x; n < a.N; n += gridDim.x, ++it) {
    mbar_wait(&full[st], ph);
    const WsRowHdr h = hdr[st];
    const float4* ws = reinterpret_cast<const float4*>(psm + (size_t)st * sbytes) + tid;
    const unsigned char* gs = psm + (size_t)st * sbytes + wbytes;
    float R = 0.f, rowdot = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if OQ_WS_OK(i) {
        const float4 w4 = ws[i * 256];
        float4 g4;
        if (G32) {
          g4 = reinterpret_cast<const float4*>(gs)[tid + i * 256];
        } else {
          const uint2 u = reinterpret_cast<const uint2*>(gs)[tid + i * 256];
          unpack_bf2(u.x, g4.x, g4.y);
          unpack_bf2(u.y, g4.z, g4.w);
        }
        float f[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float wv = f4(const_cast<float4&>(w4), j);
          const float w1 = OQ_MUL(wv, f4(colr[i], j));
          f[j] = OQ_MUL(w1, h.rmul);
          const float qa = f[j] * h.inv;
          const float qc = fminf(fmaxf(qa, h.qlo), h.qhi);
          const float rc = __fsub_rn(__fadd_rn(qc, magic), magic);     // clamp(rint(qa), -zp, qmax - zp)
          const bool inr = qc == qa;
          const float gk = f4(g4, j);
          R = fmaf(gk, rc - (inr ? qa : 0.0f), R);
          const float pw = (inr ? gk : 0.0f) * wv;
          rowdot = fmaf(pw, f4(colr[i], j), rowdot);
          f4(cacc[i], j) = fmaf(pw, h.rmul, f4(cacc[i], j));
          f4(sacc[i], j) = fmaf(wv, h.en, f4(sacc[i], j));
        }
        const float fhi = fmaxf(fmaxf(f[0], f[1]), fmaxf(f[2], f[3])), flo = fminf(fminf(f[0], f[1]), fminf(f[2], f[3]));
        if ((fhi >= h.xmax - h.tolmax) | (flo <= h.xmin + h.tolmin)) {   // rare: report the arg-max / arg-min elements to the control warp
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float wv = f4(const_cast<float4&>(w4), j);
            const float w1 = OQ_MUL(wv, f4(colr[i], j));
            const bool ismax = ROWDIV ? fabsf(f[j] - h.xmax) <= h.tolmax : f[j] == h.xmax;
            const bool ismin = ROWDIV ? fabsf(f[j] - h.xmin) <= h.tolmin : f[j] == h.xmin;
            if (ismax || ismin) {
              const int slot = atomicAdd(&tie_cnt[st], 1);
              if (slot < WS_TIE_CAP) {
                tie_k[st][slot] = ((4 * (tid + i * 256) + j) << 2) | (ismax ? 1 : 0) | (ismin ? 2 : 0);
                tie_wv[st][slot] = wv;
                tie_w1[st][slot] = w1;
              }
            }
          }
        }
      }
    }
    // one butterfly for both sums: lanes 0-15 end up with R, lanes 16-31 with rowdot
    {
      const bool up = lane & 16;
      float v = up ? rowdot : R;
      const float other = up ? R : rowdot;
      v += __shfl_xor_sync(0xffffffffu, other, 16);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((lane & 15) == 0) red[st][up ? 1 : 0][warp] = v;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);    // release: partial sums, tie list and the stage itself
    if (++st == STAGES) { st = 0; ph ^= 1u; }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if OQ_WS_OK(i) {
      const int idx = tid + i * 256;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (want_col) atomicAdd(a.dcol + 4 * idx + j, f4(cacc[i], j));
        if (want_sh) atomicAdd(a.dshift + 4 * idx + j, f4(sacc[i], j));
      }
    }
  }
#undef OQ_WS_OK
}

// rows in flight per block: the kernel is bound by bytes in flight (ncu: consumers wait on the stage's barrier), measured
// at K = 4096 (B200): 3 and 4 stages within 1 % once the control warp stopped being the bound; 6 stages (one block per SM) 1.7x slower
static int ws_stages() {
  static const int v = [] { const char* e = getenv("OQ_WQ_WS_STAGES"); return (e && atoi(e) == 3) ? 3 : 4; }();
  return v;
}
template <int NVV, int ROWDIV, bool G32, bool EXACT, int STAGES>
static bool launch_pc_bwd_ws5(const WqArgs& a, int grid, size_t smem, cudaStream_t s) {
  auto kern = wq_pc_bwd_ws<NVV, STAGES, ROWDIV, G32, EXACT>;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return false;
    attr = true;
  }
  kern<<<grid, 288, smem, s>>>(a);
  return true;
}
template <int NVV, int ROWDIV, bool G32, bool EXACT>
static bool launch_pc_bwd_ws4(const WqArgs& a, int stages, int grid, size_t smem, cudaStream_t s) {
  if (stages == 3) return launch_pc_bwd_ws5<NVV, ROWDIV, G32, EXACT, 3>(a, grid, smem, s);
  return launch_pc_bwd_ws5<NVV, ROWDIV, G32, EXACT, 4>(a, grid, smem, s);
}
template <int NVV, int ROWDIV, bool EXACT>
static bool launch_pc_bwd_ws3(const WqArgs& a, int stages, int grid, size_t smem, cudaStream_t s) {
  return a.g_is_f32 ? launch_pc_bwd_ws4<NVV, ROWDIV, true, EXACT>(a, stages, grid, smem, s)
                    : launch_pc_bwd_ws4<NVV, ROWDIV, false, EXACT>(a, stages, grid, smem, s);
}
template <int ROWDIV>
static bool launch_pc_bwd_ws2(const WqArgs& a, cudaStream_t s) {
  const int K4 = a.K / 4;
  const int nv = (K4 + 255) / 256;
  const size_t row_bytes = (size_t)a.K * (a.g_is_f32 ? 8 : 6);
  int STAGES = ws_stages();
  if (STAGES == 4 && 4 * row_bytes > 200 * 1024) STAGES = 3;
  const size_t smem = (size_t)STAGES * row_bytes;
  if (nv > 8 || smem > 200 * 1024 || a.dW) return false;
  const int bps = (int)std::max<size_t>(1, std::min<size_t>(nv <= 4 ? 2 : 1, (225 * 1024) / (smem + 12 * 1024)));
  const int grid = (int)std::min<long>(a.N, (long)oq_num_sms() * bps);
  if (nv == 1) return launch_pc_bwd_ws3<1, ROWDIV, true>(a, STAGES, grid, smem, s);
  if (nv == 2) return launch_pc_bwd_ws3<2, ROWDIV, true>(a, STAGES, grid, smem, s);
  if (nv == 3) return launch_pc_bwd_ws3<3, ROWDIV, true>(a, STAGES, grid, smem, s);
  if (nv == 4) return launch_pc_bwd_ws3<4, ROWDIV, true>(a, STAGES, grid, smem, s);
  if (nv == 8) return launch_pc_bwd_ws3<8, ROWDIV, true>(a, STAGES, grid, smem, s);
  return launch_pc_bwd_ws3<8, ROWDIV, false>(a, STAGES, grid, smem, s);
}
// row folds (multiply / divide / segment table) share the per-row table variant; "no row fold" has its own (compile time)
static bool launch_pc_bwd_ws(const WqArgs& a, cudaStream_t s) {
  if (a.nseg || (a.row && a.row_op != 0)) return launch_pc_bwd_ws2<2>(a, s);
  return launch_pc_bwd_ws2<0>(a, s);
}
