/* Host build of the PRODUCT quantiser arithmetic (omniquant_b200/csrc/oq_math.h) so the
 * exact code the CUDA kernels inline can be checked against the golden fixtures without a GPU. */
#include "../omniquant_b200/csrc/oq_math.h"

extern "C" {

int host_fake_quant(const float* W, int N, int K, int group, int n_bits, int symmetric, int recip_div,
                    const float* sig_up, const float* sig_low, float* deq, float* scale, float* zp, float* codes) {
  int gsz = group ? group : K;
  int G = (K + gsz - 1) / gsz;
  for (int n = 0; n < N; ++n)
    for (int g = 0; g < G; ++g) {
      int k0 = g * gsz, k1 = k0 + gsz < K ? k0 + gsz : K;
      float mn = INFINITY, mx = -INFINITY;
      for (int k = k0; k < k1; ++k) {
        float w = W[(long)n * K + k];
        mn = fminf(mn, w);
        mx = fmaxf(mx, w);
      }
      if (k0 + gsz > K) { mn = fminf(mn, 0.f); mx = fmaxf(mx, 0.f); }
      int item = n * G + g;
      int lwc = sig_up != 0;
      OqQParams q = oq_qparams(mn, mx, lwc ? sig_up[item] : 1.f, lwc ? sig_low[item] : 1.f, lwc, n_bits, symmetric, recip_div);
      scale[item] = q.scale;
      zp[item] = q.zp;
      for (int k = k0; k < k1; ++k) {
        float c = oq_code(W[(long)n * K + k], q);
        codes[(long)n * K + k] = c;
        deq[(long)n * K + k] = oq_dequant(c, q);
      }
    }
  return 0;
}
float host_truncate(float x, float thr) { return oq_truncate(x, thr); }
float host_fold(float w, float col, int has_col, float row, int row_op) { return oq_fold(w, col, has_col, row, row_op); }
}  // extern "C"
