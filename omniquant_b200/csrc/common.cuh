// Shared device helpers for the omniquant_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "oq_math.h"

typedef __nv_bfloat16 bf16;

// ---- error plumbing (C ABI: 0 ok, <0 bad argument, >0 cudaError_t) ----------
void oq_set_error(const char* fmt, ...);
int oq_check_launch(const char* what);
#define OQ_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      oq_set_error(__VA_ARGS__);     \
      return -1;                     \
    }                                \
  } while (0)

int oq_num_sms();

// ---- warp / block reductions -------------------------------------------------
template <typename T, typename Op>
__device__ __forceinline__ T warp_reduce(T v, Op op) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
struct OpSum {
  __device__ __forceinline__ float operator()(float a, float b) const { return a + b; }
};
struct OpMax {
  __device__ __forceinline__ float operator()(float a, float b) const { return fmaxf(a, b); }
};
struct OpMin {
  __device__ __forceinline__ float operator()(float a, float b) const { return fminf(a, b); }
};

// Block-wide reduction of NV values at once; result broadcast to all threads.
// `scratch` must hold NV*32 floats.  Safe to call repeatedly (trailing barrier).
template <int NV, typename Op>
__device__ __forceinline__ void block_reduce(float (&v)[NV], float* scratch, Op op, float identity) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_reduce(v[i], op);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) scratch[i * 32 + warp] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float x = lane < nwarps ? scratch[i * 32 + lane] : identity;
    v[i] = warp_reduce(x, op);
  }
  __syncthreads();
}

// ---- bf16 pack / unpack ---------------------------------------------------------
__device__ __forceinline__ float bf2f(bf16 x) { return __bfloat162float(x); }
__device__ __forceinline__ bf16 f2bf(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void unpack_bf2(uint32_t u, float& a, float& b) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  a = __low2float(t);
  b = __high2float(t);
}
// 4 bf16 <-> 4 floats through one 8-byte access
__device__ __forceinline__ void load_bf4(const bf16* p, float (&v)[4]) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  unpack_bf2(u.x, v[0], v[1]);
  unpack_bf2(u.y, v[2], v[3]);
}
__device__ __forceinline__ void store_bf4(bf16* p, const float (&v)[4]) {
  uint2 u;
  u.x = pack_bf2(v[0], v[1]);
  u.y = pack_bf2(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = u;
}
// 8 bf16 <-> 8 floats through one 16-byte access
__device__ __forceinline__ void load_bf8(const bf16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  unpack_bf2(u.x, v[0], v[1]);
  unpack_bf2(u.y, v[2], v[3]);
  unpack_bf2(u.z, v[4], v[5]);
  unpack_bf2(u.w, v[6], v[7]);
}
__device__ __forceinline__ void store_bf8(bf16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf2(v[0], v[1]);
  u.y = pack_bf2(v[2], v[3]);
  u.z = pack_bf2(v[4], v[5]);
  u.w = pack_bf2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ---- SwiGLU (LlamaMLP: down(silu(gate) * up), models/int_llama_layer.py:44-45); one definition shared by the
// stand-alone and the act-quant-fused kernels so both produce the same bits
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float swiglu_val(float g, float u) { return __fmul_rn(__fmul_rn(g, sigmoid_fast(g)), u); }
__device__ __forceinline__ void swiglu_grad(float g, float u, float d, float& dg, float& du) {
  const float s = sigmoid_fast(g);
  dg = __fmul_rn(__fmul_rn(__fmul_rn(d, u), s), __fmaf_rn(g, 1.0f - s, 1.0f));
  du = __fmul_rn(__fmul_rn(d, g), s);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ---- s8 side output of the per-token / per-channel quantisers: the operands of oq_gemm_s8 -------------------------
struct S8Side {
  int8_t* codes; long ld;   // [rows, C] signed codes (code - z0), row stride ld bytes
  float* meta;              // [3, rows]: scale | t = scale * sum_k (code - z0) | beta = (z0 - zp) * scale
  long rows;
};
// (code - z0) is an integer-valued float with |.| <= 255: adding 1.5 * 2^23 leaves the two's-complement integer in the low
// mantissa bits, so the int8 is the low byte of the float's bit pattern (1 FADD instead of a convert + mask + shift);
// four such bytes are gathered with two PRMT + one PRMT, and their sum is ONE dp4a against 0x01010101.
__device__ __forceinline__ uint32_t s8_pack4(float d0, float d1, float d2, float d3) {
  const float magic = 12582912.0f;
  const uint32_t b0 = __float_as_uint(d0 + magic), b1 = __float_as_uint(d1 + magic);
  const uint32_t b2 = __float_as_uint(d2 + magic), b3 = __float_as_uint(d3 + magic);
  return __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
}
// 8 consecutive codes -> 8 bytes (one 8-byte store); adds sum (code - z0) to the integer accumulator `isum`
__device__ __forceinline__ void s8_store8(int8_t* p, const float (&c)[8], float z0, int& isum) {
  const uint32_t w0 = s8_pack4(c[0] - z0, c[1] - z0, c[2] - z0, c[3] - z0);
  const uint32_t w1 = s8_pack4(c[4] - z0, c[5] - z0, c[6] - z0, c[7] - z0);
  isum = __dp4a((int)w0, 0x01010101, isum);
  isum = __dp4a((int)w1, 0x01010101, isum);
  *reinterpret_cast<uint2*>(p) = make_uint2(w0, w1);
}
__device__ __forceinline__ void s8_store4(int8_t* p, const float (&c)[4], float z0, int& isum) {
  const uint32_t w = s8_pack4(c[0] - z0, c[1] - z0, c[2] - z0, c[3] - z0);
  isum = __dp4a((int)w, 0x01010101, isum);
  *reinterpret_cast<uint32_t*>(p) = w;
}
// a degenerate row (constant: scale 0 / zero point NaN) dequantises to NaN in the fake-quant path; the integer path
// reports a NaN scale so that the GEMM output row is NaN as well
__device__ __forceinline__ float s8_scale_or_nan(const OqQParams& q) {
  return (isfinite(q.zp) && isfinite(q.scale)) ? q.scale : __int_as_float(0x7fc00000);
}
__device__ __forceinline__ void s8_write_meta(const S8Side& s, long r, const OqQParams& q, float z0, float csum) {
  s.meta[r] = s8_scale_or_nan(q);
  s.meta[s.rows + r] = q.scale * csum;
  s.meta[2 * s.rows + r] = (z0 - q.zp) * q.scale;
}

static inline int ceil_div_i(long a, long b) { return (int)((a + b - 1) / b); }
