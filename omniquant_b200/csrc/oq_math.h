// Scalar arithmetic of the uniform affine fake-quantiser, shared by the CUDA
// kernels (device) and the CPU self-check of the product code (host).
//
// Follows UniformAffineQuantizer.per_token_dynamic_calibration and fake_quant
// (reference quantize/quantizer.py:122-147 and :84-105) one ATen op per
// rounding step, so integer codes / scales / zero points are bit-exact:
//   - every +,-,*,/ is an individually rounded fp32 op (no FMA contraction),
//   - rounding is half-to-even (rintf),
//   - clamp keeps NaN like torch.clamp (a constant group gives scale 0 -> NaN,
//     exactly as the reference does; its asymmetric scale clamp is dead code,
//     quantizer.py:144-145),
//   - `tensor / python_scalar` is a true division on CPU but a multiply by the
//     fp32 reciprocal in ATen's CUDA kernel; `recip_div` selects which.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define OQ_HD __host__ __device__ __forceinline__
#define OQ_MUL(a, b) __fmul_rn((a), (b))
#define OQ_ADD(a, b) __fadd_rn((a), (b))
#define OQ_SUB(a, b) __fsub_rn((a), (b))
#define OQ_DIV(a, b) __fdiv_rn((a), (b))
#elif defined(__CUDACC__)
#define OQ_HD __host__ __device__ __forceinline__
#define OQ_MUL(a, b) ((a) * (b))
#define OQ_ADD(a, b) ((a) + (b))
#define OQ_SUB(a, b) ((a) - (b))
#define OQ_DIV(a, b) ((a) / (b))
#else
#define OQ_HD static inline
#define OQ_MUL(a, b) ((a) * (b))
#define OQ_ADD(a, b) ((a) + (b))
#define OQ_SUB(a, b) ((a) - (b))
#define OQ_DIV(a, b) ((a) / (b))
#endif

struct OqQParams {
  float scale;  // quantizer.py:139/:145
  float zp;     // rounded zero point, quantizer.py:147
  float qmax;   // 2^n - 1
};

// torch.clamp semantics (NaN propagates)
OQ_HD float oq_clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

OQ_HD float oq_sigmoidf(float x) {
  // ATen sigmoid for float: 1 / (1 + exp(-x)) in fp32
  return OQ_DIV(1.0f, OQ_ADD(1.0f, expf(-x)));
}

// xmin/xmax: raw group extrema; su/sl: sigmoid(upbound_factor), sigmoid(lowbound_factor)
// (pass lwc=0 to skip the clip).  quantizer.py:131-147.
OQ_HD OqQParams oq_qparams(float xmin, float xmax, float su, float sl, int lwc, int n_bits, int symmetric,
                           int recip_div) {
  OqQParams q;
  q.qmax = (float)((1 << n_bits) - 1);
  if (lwc) {
    xmax = OQ_MUL(su, xmax);
    xmin = OQ_MUL(sl, xmin);
  }
  if (symmetric) {
    float amax = fmaxf(fabsf(xmax), fabsf(xmin));
    float lv = (float)((1 << (n_bits - 1)) - 1);
    float s = recip_div ? OQ_MUL(amax, OQ_DIV(1.0f, lv)) : OQ_DIV(amax, lv);
    q.scale = oq_clampf(s, 1e-5f, 1e4f);
    q.zp = lv;
  } else {
    float range = OQ_SUB(xmax, xmin);
    q.scale = recip_div ? OQ_MUL(range, OQ_DIV(1.0f, q.qmax)) : OQ_DIV(range, q.qmax);
    q.zp = OQ_DIV(-xmin, q.scale);
  }
  q.zp = rintf(oq_clampf(q.zp, -1e4f, 1e4f));
  return q;
}

// integer code before the clamp: round(x/scale) + zp   (quantizer.py:93-95)
OQ_HD float oq_code_unclamped(float x, const OqQParams& q) { return OQ_ADD(rintf(OQ_DIV(x, q.scale)), q.zp); }

// clamped code (quantizer.py:96)
OQ_HD float oq_code(float x, const OqQParams& q) { return oq_clampf(oq_code_unclamped(x, q), 0.0f, q.qmax); }

// dequantised value from a code (quantizer.py:97-100)
OQ_HD float oq_dequant(float code, const OqQParams& q) { return OQ_MUL(OQ_SUB(code, q.zp), q.scale); }

OQ_HD float oq_fake_quant(float x, const OqQParams& q) { return oq_dequant(oq_code(x, q), q); }

// Exact-integer operand of oq_gemm_s8: the stored signed code is (code - z0) with z0 = the zero point clamped so that
// every code - z0 fits in int8 (for n <= 7 bits: z0 = clamp(zp, 0, 2^n - 1); for 8 bits z0 = 128).  A zero point outside
// that range (a row that does not straddle zero) leaves the remainder beta = (z0 - zp) * scale to the GEMM epilogue.
OQ_HD float oq_s8_z0(float zp, float qmax) {
  const float lo = fmaxf(0.0f, qmax - 127.0f), hi = fminf(qmax, 128.0f);
  return fminf(fmaxf(zp, lo), hi);     // NaN zp -> lo (the row's scale is reported as NaN instead)
}

// LET fold of one weight element: (W * col) {none | * row | / row}
// (models/transformation.py:41,51,60,66,68).  row_op: 0 none, 1 mul, 2 div.
OQ_HD float oq_fold(float w, float col, int has_col, float row, int row_op) {
  if (has_col) w = OQ_MUL(w, col);
  if (row_op == 1) w = OQ_MUL(w, row);
  else if (row_op == 2) w = OQ_DIV(w, row);
  return w;
}

// truncate_number, models/transformation.py:5-20 (|x| < thr -> sign(x)*thr; sign(0)=0)
OQ_HD float oq_truncate(float x, float thr) {
  if (fabsf(x) < thr) {
    float s = (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f);
    return OQ_MUL(s, thr);
  }
  return x;
}

// Gradient of the loss w.r.t. the clipped extrema xmax' = su*xmax, xmin' = sl*xmin given
// R = dL/dscale.  Asymmetric: scale = (xmax'-xmin')/Q.  Symmetric: scale =
// clamp(max(|xmax'|,|xmin'|)/(2^(n-1)-1), 1e-5, 1e4); torch.max splits ties evenly and the
// clamp passes gradient inside its bounds.  The rounded zero point carries no gradient
// (quantizer.py:147).
OQ_HD void oq_scale_grad(float R, float xmin, float xmax, float su, float sl, int lwc, int n_bits, int symmetric,
                         int recip_div, float* gmaxp, float* gminp) {
  if (symmetric) {
    const float xp = lwc ? OQ_MUL(su, xmax) : xmax, xm = lwc ? OQ_MUL(sl, xmin) : xmin;
    const float lv = (float)((1 << (n_bits - 1)) - 1);
    const float amax = fmaxf(fabsf(xp), fabsf(xm));
    const float sraw = recip_div ? OQ_MUL(amax, OQ_DIV(1.0f, lv)) : OQ_DIV(amax, lv);
    const float pass = (sraw >= 1e-5f && sraw <= 1e4f) ? R / lv : 0.0f;
    const float wa = fabsf(xp) > fabsf(xm) ? 1.0f : (fabsf(xp) == fabsf(xm) ? 0.5f : 0.0f);
    const float sgp = xp > 0.f ? 1.f : (xp < 0.f ? -1.f : 0.f), sgm = xm > 0.f ? 1.f : (xm < 0.f ? -1.f : 0.f);
    *gmaxp = pass * wa * sgp;
    *gminp = pass * (1.0f - wa) * sgm;
  } else {
    const float Q = (float)((1 << n_bits) - 1);
    *gmaxp = R / Q;
    *gminp = -R / Q;
  }
}

// rint(x / scale) with the division replaced by a multiply in the common case, bit-identical to
// rintf(x / scale): the product x*inv_scale is within 1.8e-7*|q| of the correctly rounded quotient, so the
// two can only round to different integers when the product sits that close to a half-integer -- then (and
// for NaN/Inf) the exact IEEE division decides.  `inv_ok` = isnormal(inv_scale), evaluated once per group.
OQ_HD float oq_rint_div(float x, float scale, float inv_scale, int inv_ok) {
  if (!inv_ok) return rintf(OQ_DIV(x, scale));
  const float qa = OQ_MUL(x, inv_scale);
  float r = rintf(qa);
  const float d = fabsf(qa - r);
  if (!(fabsf(d - 0.5f) > fabsf(qa) * 2.4e-7f + 1e-30f)) r = rintf(OQ_DIV(x, scale));
  return r;
}
OQ_HD int oq_inv_ok(float inv_scale) {
  const float a = fabsf(inv_scale);
  return (a >= 1.1754944e-38f) && (a <= 3.4028235e38f);
}
