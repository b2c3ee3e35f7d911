// GPTQ-style packing of a calibrated linear layer: replaces the reference's
//   q_linear = qlinear_cuda.QuantLinear(wbits, group_size, in, out, bias); q_linear.pack(module, scales, zeros)
// (quantize/omniquant.py:255-265; the packer itself is the third-party AutoGPTQ `QuantLinear.pack`, absent from the
// reference tree -- oracle/gptq_pack.py restates the published upstream algorithm and says what is unpinned).
//
//   intweight[k, n] = rint((W[n, k] + scale[n, g] * zero[n, g]) / float(half(scale[n, g]))),  g = k / group
//   qweight [K/32*bits, N] int32 : the codes of 32 consecutive k form a little-endian bit stream of 32*bits bits
//                                  (bits 2/4/8: 32/bits codes per word; bits 3: AutoGPTQ's 10+1|10+1|10 split is the
//                                  same stream), one column per output feature n
//   qzeros  [G, N/32*bits] int32 : (zero - 1), same stream along n
//   scales  [G, N] fp16, g_idx [K] int32
// Byte work bound by HBM: 4 B read per weight, bits/8 B written.  A block converts a 64(n) x 128(k) tile (rows read
// as 512 contiguous bytes), stages the codes in shared memory and writes words with n fastest (256-byte segments).
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"
#include "oq_math.h"

namespace {

constexpr int PT_N = 64, PT_K = 128;

__device__ __forceinline__ uint32_t stream_word(const uint8_t* codes, int stride, int bits, int wi) {
  // word `wi` of the little-endian stream of `bits`-bit fields codes[j * stride]
  const int b0 = 32 * wi;
  int j = b0 / bits;
  uint32_t w = 0;
  for (int pos = j * bits; pos < b0 + 32; pos += bits, ++j) {
    const uint32_t c = codes[j * stride];
    const int sh = pos - b0;
    w |= sh >= 0 ? (c << sh) : (c >> (-sh));
  }
  return w;
}

__global__ void __launch_bounds__(256) gptq_pack_weight(const float* __restrict__ W, const float* __restrict__ scales,
                                                        const float* __restrict__ zeros, int N, int K, int group, int G,
                                                        int bits, int32_t* __restrict__ qweight) {
  __shared__ uint8_t code[PT_K][PT_N + 4];     // [k][n]: the word builder walks k at fixed n
  const int n0 = blockIdx.x * PT_N, k0 = blockIdx.y * PT_K;
  const uint32_t mask = (1u << bits) - 1u;
  // 64 rows x 32 float4 per row
  for (int i = threadIdx.x; i < PT_N * (PT_K / 4); i += 256) {
    const int r = i / (PT_K / 4), c4 = i % (PT_K / 4);
    const int n = n0 + r, k = k0 + 4 * c4;
    if (n < N && k < K) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(W + (long)n * K + k));
      const float* wv = reinterpret_cast<const float*>(&w4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int g = (k + j) / group;
        const float s = scales[(long)n * G + g], z = zeros[(long)n * G + g];
        const float den = __half2float(__float2half_rn(s));
        const float q = rintf(OQ_DIV(OQ_ADD(wv[j], OQ_MUL(z, s)), den));
        code[4 * c4 + j][r] = (uint8_t)((uint32_t)(int)q & mask);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) code[4 * c4 + j][r] = 0;
    }
  }
  __syncthreads();
  const int words = PT_K * bits / 32;          // words per column in this tile
  for (int i = threadIdx.x; i < words * PT_N; i += 256) {
    const int wi = i / PT_N, r = i % PT_N;
    const int n = n0 + r;
    const long row = (long)k0 * bits / 32 + wi;
    if (n < N && row < (long)K * bits / 32) qweight[row * N + n] = (int32_t)stream_word(&code[0][r], PT_N + 4, bits, wi);
  }
}

__global__ void gptq_pack_meta(const float* __restrict__ scales, const float* __restrict__ zeros, int N, int K, int group,
                               int G, int bits, int32_t* __restrict__ qzeros, __half* __restrict__ scales_h,
                               int32_t* __restrict__ g_idx) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t mask = (1u << bits) - 1u;
  const int zw = N * bits / 32;                // qzeros words per group
  if (t < (long)G * zw) {
    const int g = (int)(t / zw), wi = (int)(t % zw);
    const int b0 = 32 * wi;
    int j = b0 / bits;
    uint32_t w = 0;
    for (int pos = j * bits; pos < b0 + 32; pos += bits, ++j) {
      const uint32_t c = (uint32_t)((int)zeros[(long)j * G + g] - 1) & mask;
      const int sh = pos - b0;
      w |= sh >= 0 ? (c << sh) : (c >> (-sh));
    }
    qzeros[t] = (int32_t)w;
  }
  if (t < (long)G * N) {
    const int g = (int)(t / N), n = (int)(t % N);
    scales_h[t] = __float2half_rn(scales[(long)n * G + g]);
  }
  if (t < K) g_idx[t] = (int32_t)(t / group);
}

}  // namespace

extern "C" int oq_gptq_pack(const float* W, const float* scales, const float* zeros, int N, int K, int group_size, int n_bits,
                            int32_t* qweight, int32_t* qzeros, void* scales_f16, int32_t* g_idx, cudaStream_t stream) {
  OQ_REQUIRE(W && scales && zeros && qweight && qzeros && scales_f16 && g_idx, "oq_gptq_pack: null argument");
  OQ_REQUIRE(n_bits == 2 || n_bits == 3 || n_bits == 4 || n_bits == 8, "oq_gptq_pack: bits %d not in {2,3,4,8}", n_bits);
  OQ_REQUIRE(N > 0 && K > 0 && N % 32 == 0 && K % 32 == 0, "oq_gptq_pack: N=%d and K=%d must be multiples of 32", N, K);
  const int group = group_size > 0 ? group_size : K;
  OQ_REQUIRE(K % group == 0, "oq_gptq_pack: K %% group_size != 0");
  OQ_REQUIRE((uintptr_t)W % 16 == 0, "oq_gptq_pack: W must be 16-byte aligned");
  const int G = K / group;
  gptq_pack_weight<<<dim3((N + PT_N - 1) / PT_N, (K + PT_K - 1) / PT_K), 256, 0, stream>>>(W, scales, zeros, N, K, group, G,
                                                                                            n_bits, qweight);
  int rc = oq_check_launch("oq_gptq_pack");
  if (rc) return rc;
  const long work = std::max<long>((long)G * N, K);
  gptq_pack_meta<<<(unsigned)((work + 255) / 256), 256, 0, stream>>>(scales, zeros, N, K, group, G, n_bits, qzeros,
                                                                     (__half*)scales_f16, g_idx);
  return oq_check_launch("oq_gptq_pack");
}
