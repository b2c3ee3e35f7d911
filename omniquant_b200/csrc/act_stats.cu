// Per-channel activation statistics for the LET initialisation (reference generate_act_scale_shift.py:25-94):
//   act_scales[name] = max over samples of max_tokens |x|            (:30-37)
//   act_shifts[name] = (max+min)/2 of the first sample, then 0.99*prev + 0.01*(max+min)/2 per sample   (:64-72)
// for the input of a linear layer.  Here the statistics are an epilogue of the full-precision pass of the block
// calibration (the FP pass sees exactly the activations of the unquantised model), so `--let` needs no separate run.
//
// oq_act_stats      : one sample [rows, C] bf16 -> per-channel abs-max and mid-point (HBM-bound column reduction:
//                     2 B read per activation; a warp reads 512 contiguous bytes of a row, 8 channels per lane).
// oq_act_stats_fold : the per-sample vectors of all samples, in sample order -> running max and the EMA, one thread per
//                     channel with the reference's individually rounded fp32 operations (bit-exact).
#include <algorithm>

#include "common.cuh"
#include "oq_math.h"

namespace {

constexpr int ST_COLS = 256;   // channels per block: 32 lanes x 8 bf16
constexpr int ST_WARPS = 8;

__global__ void __launch_bounds__(256) act_stats_partial(const bf16* __restrict__ x, int rows, int C, long ld, int rows_per_chunk, int vec,
                                                         float* __restrict__ ws_max, float* __restrict__ ws_min) {
  __shared__ float smax[ST_WARPS][ST_COLS], smin[ST_WARPS][ST_COLS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * ST_COLS + lane * 8;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
  float mx[8], mn[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { mx[j] = -INFINITY; mn[j] = INFINITY; }
  if (vec && c0 + 8 <= C) {
    for (int r = r0 + warp; r < r1; r += ST_WARPS) {
      float v[8];
      load_bf8(x + (long)r * ld + c0, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) { mx[j] = fmaxf(mx[j], v[j]); mn[j] = fminf(mn[j], v[j]); }
    }
  } else {
    for (int r = r0 + warp; r < r1; r += ST_WARPS) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (c0 + j < C) {
          const float v = bf2f(x[(long)r * ld + c0 + j]);
          mx[j] = fmaxf(mx[j], v);
          mn[j] = fminf(mn[j], v);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { smax[warp][lane * 8 + j] = mx[j]; smin[warp][lane * 8 + j] = mn[j]; }
  __syncthreads();
  const int c = blockIdx.x * ST_COLS + threadIdx.x;
  if (c < C) {
    float a = -INFINITY, b = INFINITY;
#pragma unroll
    for (int w = 0; w < ST_WARPS; ++w) { a = fmaxf(a, smax[w][threadIdx.x]); b = fminf(b, smin[w][threadIdx.x]); }
    ws_max[(long)blockIdx.y * C + c] = a;
    ws_min[(long)blockIdx.y * C + c] = b;
  }
}

__global__ void act_stats_final(const float* __restrict__ ws_max, const float* __restrict__ ws_min, int chunks, int C,
                                float* __restrict__ absmax, float* __restrict__ mid) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = -INFINITY, b = INFINITY;
  for (int k = 0; k < chunks; ++k) { a = fmaxf(a, ws_max[(long)k * C + c]); b = fminf(b, ws_min[(long)k * C + c]); }
  if (absmax) absmax[c] = fmaxf(fabsf(a), fabsf(b));            // max |x| = max(|max|, |min|)
  if (mid) mid[c] = OQ_DIV(OQ_ADD(a, b), 2.0f);                 // generate_act_scale_shift.py:70,72
}

__global__ void act_stats_fold(const float* __restrict__ absmax, const float* __restrict__ mid, int n, int C, int first,
                               float* __restrict__ scale_io, float* __restrict__ shift_io) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (scale_io && absmax) {
    float s = first ? absmax[c] : scale_io[c];
    for (int i = first ? 1 : 0; i < n; ++i) s = fmaxf(s, absmax[(long)i * C + c]);      // :34-35
    scale_io[c] = s;
  }
  if (shift_io && mid) {
    float s = first ? mid[c] : shift_io[c];
    for (int i = first ? 1 : 0; i < n; ++i) s = OQ_ADD(OQ_MUL(0.99f, s), OQ_MUL(0.01f, mid[(long)i * C + c]));   // :69-70
    shift_io[c] = s;
  }
}

}  // namespace

static long stats_ws_bytes(int rows, int C) {
  const int strips = (C + ST_COLS - 1) / ST_COLS;
  const int chunks = std::max(1, std::min((2 * oq_num_sms() + strips - 1) / strips, (rows + 31) / 32));
  return (long)2 * chunks * C * sizeof(float);
}

extern "C" int oq_act_stats_workspace(int rows, int C, long* bytes_out) {
  OQ_REQUIRE(rows > 0 && C > 0 && bytes_out, "oq_act_stats_workspace: bad arguments");
  *bytes_out = stats_ws_bytes(rows, C);
  return 0;
}

extern "C" int oq_act_stats(const void* x, int rows, int C, long ld, float* absmax, float* mid, float* workspace,
                            long workspace_bytes, cudaStream_t stream) {
  OQ_REQUIRE(x && rows > 0 && C > 0 && ld >= C, "oq_act_stats: bad shape rows=%d C=%d ld=%ld", rows, C, ld);
  OQ_REQUIRE(absmax || mid, "oq_act_stats: no output requested");
  const long need = stats_ws_bytes(rows, C);
  OQ_REQUIRE(workspace && workspace_bytes >= need, "oq_act_stats: workspace of %ld bytes needed", need);
  const bool vec = (ld % 8 == 0) && ((uintptr_t)x % 16 == 0);
  const int strips = (C + ST_COLS - 1) / ST_COLS;
  const int chunks = (int)(need / ((long)2 * C * sizeof(float)));
  const int rpc = (rows + chunks - 1) / chunks;
  float* ws_max = workspace;
  float* ws_min = workspace + (long)chunks * C;
  act_stats_partial<<<dim3(strips, chunks), 256, 0, stream>>>((const bf16*)x, rows, C, ld, rpc, vec ? 1 : 0, ws_max, ws_min);
  int rc = oq_check_launch("oq_act_stats");
  if (rc) return rc;
  act_stats_final<<<(C + 255) / 256, 256, 0, stream>>>(ws_max, ws_min, chunks, C, absmax, mid);
  return oq_check_launch("oq_act_stats");
}

extern "C" int oq_act_stats_fold(const float* absmax, const float* mid, int n, int C, int first, float* scale_io,
                                 float* shift_io, cudaStream_t stream) {
  OQ_REQUIRE(n > 0 && C > 0, "oq_act_stats_fold: bad shape n=%d C=%d", n, C);
  OQ_REQUIRE((scale_io && absmax) || (shift_io && mid), "oq_act_stats_fold: nothing to do");
  act_stats_fold<<<(C + 255) / 256, 256, 0, stream>>>(absmax, mid, n, C, first, scale_io, shift_io);
  return oq_check_launch("oq_act_stats_fold");
}
