/* omniquant_b200 -- C ABI of the B200 (sm_100a) calibration hot path.
 *
 * Drop-in boundary.  The reference (SuperVan-Young/OmniQuant) is
 * pure PyTorch and has no FFI of its own: its operator API is the nn.Module
 * surface.  These entry points are what that surface binds to here -- each
 * comment cites the reference code the call replaces (paths relative to the
 * reference repo).  The Python mirror of the modules lives in
 * omniquant_b200/{quantize,models}/ and calls only these functions for device
 * compute (ctypes; see INTEGRATION.md for the binding a maintainer would add).
 *
 * Conventions
 *   - every function returns int: 0 ok, <0 invalid argument, >0 a cudaError_t;
 *     oq_last_error() returns a thread-local message for the last failure;
 *   - the caller owns all memory; every pointer is a DEVICE pointer unless
 *     named *_host; the library never allocates, frees or synchronises;
 *   - all work is enqueued on the explicit `stream` (last argument);
 *   - tensors are row-major; "ld" arguments are leading dimensions in ELEMENTS;
 *   - dtypes are fixed per argument: `void*` activations / operands are bf16,
 *     `float*` are fp32 master weights, parameters, statistics and reductions;
 *   - no C++ exceptions, no torch types.
 */
#ifndef OMNIQUANT_B200_H_
#define OMNIQUANT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* oq_stream_t; /* == cudaStream_t */

int oq_version(void);
const char* oq_last_error(void);
int oq_sm_count(void);

/* ---- K0: weight path -----------------------------------------------------
 * LET fold + LWC clip + per-channel / per-group fake-quant of one weight.
 *   W' = (W * col_scale[k]) {row_op: 0 none | 1 * row_scale[n] | 2 / row_scale[n]}
 *        models/transformation.py:41 (smooth_ln_fcs), :51,:60 (smooth_fc_fc), :66,:68 (smooth_q_k)
 *   matvec_out[n] = sum_k W[n,k]*shift[k]   (raw W; transformation.py:38-40,:57-59)
 *   group extrema -> sigmoid(up/low) clip -> scale, rounded zero point
 *        quantize/quantizer.py:122-147 (UniformAffineQuantizer.per_token_dynamic_calibration)
 *   codes = clamp(round(W'/s)+z, 0, 2^n-1); W_q = (codes-z)*s   quantizer.py:84-105 (fake_quant)
 * group_size 0 = per output channel.  up/low: LWC factors [N*ceil(K/group)] or both NULL.
 * recip_div: 0 = CPU ATen semantics for `range / (2^n-1)` (true division), 1 = CUDA ATen
 * semantics (multiply by the fp32 reciprocal).  Any of the outputs may be NULL.  n_bits >= 16: fold only (the quantiser
 * is the identity, quantize/quantizer.py:109) -- this mode also implements smooth_*_{temporary,inplace} on their own.
 * codes_s8 / s8_t / s8_beta (all or none; per-channel, needs scale_out): the exact-integer operand of oq_gemm_s8, see there.
 * Row segments (n_seg in 1..3, HOST arrays; row_scale must be NULL): one launch over several weights stacked along N that
 * share col_scale / shift but differ in the row fold -- rows [seg_end[s-1], seg_end[s]) use the DEVICE vector seg_row[s]
 * (indexed from the segment's first row; NULL = no row fold) with operation seg_op[s].  This is q | k | v of one block:
 * smooth_ln_fcs on all three, then smooth_q_k on q (/) and k (*), smooth_fc_fc on v (/) (models/int_llama_layer.py:285-292).
 * Replaces models/int_llama_layer.py:285-304 (smooth_and_quant_temporary) per weight. */
int oq_weight_quant_fwd(const float* W, int N, int K, const float* col_scale, const float* row_scale, int row_op,
                        const float* shift, float* matvec_out, int n_bits, int symmetric, int group_size,
                        const float* up, const float* low, int recip_div, void* wq_bf16, long ld_bf16, float* wq_f32,
                        uint8_t* codes, float* scale_out, float* zp_out, float* xmin_out, float* xmax_out,
                        int8_t* codes_s8, float* s8_t, float* s8_beta, int n_seg, const int* seg_end_host,
                        const float* const* seg_row_host, const int* seg_op_host, oq_stream_t stream);

/* Backward of the above (autograd of quantizer.py:84-147 + transformation.py folds).
 * grad_wq: dL/dW_q [N,K] (bf16, or fp32 when grad_is_f32).  xmin/xmax: raw group extrema saved by the
 * forward; scale_saved/zp_saved (optional, both or neither): its scale_out/zp_out, which spares the backward
 * the per-row sigmoid + division chain.  Writes d_up/d_low [N*G]; writes drow[n] = sum_k dW'[n,k]*(W*col)[n,k]; atomically ADDS
 * dcol_acc[k] += sum_n dW'[n,k]*rowfac[n]*W[n,k] and dshift_acc[k] += sum_n W[n,k]*e_row[n]
 * (caller zeroes them once per step); optionally writes the full dL/dW' (fp32 [N,K]). */
int oq_weight_quant_bwd(const float* W, int N, int K, const float* col_scale, const float* row_scale, int row_op,
                        int n_bits, int symmetric, int group_size, const float* up, const float* low, int recip_div,
                        const float* xmin, const float* xmax, const float* scale_saved, const float* zp_saved,
                        const void* grad_wq, int grad_is_f32, long ld_grad,
                        const float* e_row, float* d_up, float* d_low, float* dcol_acc, float* drow, float* dshift_acc,
                        float* dW_full, int n_seg, const int* seg_end_host, const float* const* seg_row_host,
                        const int* seg_op_host, oq_stream_t stream);

/* ---- K4: per-token dynamic activation fake-quant ---------------------------
 * UniformAffineQuantizer(dynamic_method="per_token") forward, quantizer.py:108-147,:84-105, as called by
 * QuantLinear.forward (quantize/int_linear.py:59-60) and QuantMatMul.quant_x1/x2 (quantize/int_matmul.py:31-39).
 * Row r lives at x + (r / rows_inner)*ld_outer + (r % rows_inner)*ld_inner, C contiguous bf16 elements.
 * rope_cos/sin ([seq_len, C] fp32, or NULL): rotary embedding applied before quantisation
 * (models/int_llama_layer.py:124-125), position = (r / rows_inner) % seq_len; needs C in {64,128}.
 * Only the inner rows [0, rope_inner) are rotated (rope_inner <= 0: all of them) -- one launch covers the fused
 * [q heads | k heads | v heads] buffer of a token, V un-rotated (:162).  n_bits >= 16 with rope = rotary only
 * (un-rotated rows are copied).
 * codes_s8 [rows, C] (row stride ld_s8 bytes) + s8_meta [3, rows] (both or neither; plain 2-D rows without rope): the
 * exact-integer operand of oq_gemm_s8 -- signed codes (code - z0), then scale | t | beta per row, see oq_gemm_s8. */
int oq_act_quant_fwd(const void* x, long rows, int C, int rows_inner, long ld_outer, long ld_inner, int n_bits,
                     int symmetric, int recip_div, const float* rope_cos, const float* rope_sin, int seq_len,
                     int rope_inner, void* y, long ldy_outer, long ldy_inner, float* scale_out, float* zp_out,
                     int8_t* codes_s8, long ld_s8, float* s8_meta, oq_stream_t stream);
/* Backward: gx = d/dx [ fake_quant(rope(x)) ] . gy   (x is the forward INPUT, re-quantised on the fly). */
int oq_act_quant_bwd(const void* x, long rows, int C, int rows_inner, long ld_outer, long ld_inner, int n_bits,
                     int symmetric, int recip_div, const float* rope_cos, const float* rope_sin, int seq_len,
                     int rope_inner, const void* gy, long ldg_outer, long ldg_inner, void* gx, long ldgx_outer,
                     long ldgx_inner, oq_stream_t stream);

/* SwiGLU fused with the per-token quantiser of down_proj's input: mq = fake_quant(bf16(silu(gate) * up)) for
 * gu = [gate | up] [rows, 2F] (LlamaMLP, models/int_llama_layer.py:44-45, feeding QuantLinear.act_quantizer,
 * quantize/int_linear.py:59-60).  The un-quantised product is never written; the backward re-derives it from gu and
 * returns dgu = d/dgu of the same composition.  Bit-identical to oq_swiglu_fwd + oq_act_quant_fwd (and their
 * backwards) run one after the other. */
int oq_swiglu_act_quant_fwd(const void* gu, long ldgu, long rows, int F, int n_bits, int symmetric, int recip_div,
                            void* mq, long ldmq, int8_t* codes_s8, long ld_s8, float* s8_meta, oq_stream_t stream);
int oq_swiglu_act_quant_bwd(const void* gu, long ldgu, long rows, int F, int n_bits, int symmetric, int recip_div,
                            const void* gy, long ldgy, void* dgu, long lddgu, oq_stream_t stream);

/* Exact-integer forward contraction of QuantLinear when BOTH operands are quantised (W4A4 / W6A6,
 * quantize/int_linear.py:48-65 with quantize/quantizer.py:93-100): A [M,K], B [N,K] hold signed 8-bit codes
 * (code - z0) K-major, the K sum is exact (tcgen05.mma kind::i8, s32 accumulators),
 *   C[m,n] = alpha * ( row_scale[m]*col_scale[n]*sum_k A[m,k]*B[n,k]
 *                      + row_t[m]*col_beta[n] + row_beta[m]*col_t[n] + K*row_beta[m]*col_beta[n] ) (+ bias[n]) (+ residual[m,n]).
 * z0 is the zero point clamped into the int8-representable range (= the zero point itself whenever the row / channel
 * straddles zero); the four correction vectors (all or none) carry the remainder: beta = (z0 - z)*scale,
 * t = scale * sum_k (code - z0).  The quantiser kernels emit codes, scale, t and beta as their "s8 side output"
 * (oq_act_quant_fwd, oq_norm_act_quant_fwd, oq_swiglu_act_quant_fwd: s8_meta = [scale | t | beta], each [rows];
 * oq_weight_quant_fwd: scale_out, s8_t, s8_beta).  lda / ldb in bytes (= elements), multiples of 16. */
int oq_gemm_s8(const void* A, long lda, const void* B, long ldb, void* C, int c_dtype, long ldc, int M, int N, int K,
               const float* row_scale, const float* col_scale, const float* row_t, const float* row_beta,
               const float* col_t, const float* col_beta, float alpha, const float* bias, const void* residual, long ldr,
               oq_stream_t stream);

/* ---- K7: norms with temporary (folded) weight / bias ------------------------
 * kind 0: OmniLlamaRMSNorm.forward quantize/omni_norm.py:52-63; kind 1: OmniLayerNorm.forward :26-33.
 * w/b are the temp_weight/temp_bias of the step (b may be NULL). */
int oq_norm_fwd(const void* x, long rows, int C, long ldx, const float* w, const float* b, float eps, int kind, void* y,
                long ldy, float* mean_out, float* rstd_out, oq_stream_t stream);
/* oq_norm_fwd plus yq = per-token fake-quant of the (bf16-rounded) norm output in the same pass: the norm feeding
 * QuantLinear.act_quantizer of q/k/v or gate/up (quantize/int_linear.py:59-60).  y may be NULL.  Bit-identical to
 * oq_norm_fwd followed by oq_act_quant_fwd. */
int oq_norm_act_quant_fwd(const void* x, long rows, int C, long ldx, const float* w, const float* b, float eps, int kind,
                          void* y, long ldy, float* mean_out, float* rstd_out, int n_bits, int symmetric, int recip_div,
                          void* yq, long ldyq, int8_t* codes_s8, long ld_s8, float* s8_meta, oq_stream_t stream);
/* dx = norm'(x).dy (+ dx_add);  dw_acc[c] += sum_r dy*xhat, db_acc[c] += sum_r dy (atomic adds; NULL to skip). */
int oq_norm_bwd(const void* x, long rows, int C, long ldx, const float* w, int kind, const float* mean,
                const float* rstd, const void* dy, long lddy, void* dx, long lddx, const void* dx_add, long ldadd,
                float* dw_acc, float* db_acc, oq_stream_t stream);

/* ---- K1/K2/K3/K5: tcgen05 bf16 GEMM ----------------------------------------
 * C[b] = alpha * A[b] * B[b]^T (+ bias[n]) (+ residual[b]),  fp32 accumulation in TMEM.
 *   a_mn=0: A[m,k] at A + b*strideA + m*lda + k     a_mn=1: A[m,k] at A + b*strideA + k*lda + m
 *   b_mn=0: B[n,k] at B + b*strideB + n*ldb + k     b_mn=1: B[n,k] at B + b*strideB + k*ldb + n
 * c_dtype: 0 bf16, 1 fp32, 2 fp16 (the attention scores, as the reference's fp16 autocast stores them).  causal: 0 none; 1 skip output tiles strictly above the diagonal (QK^T, dP);
 * 2 reduce only k < m0+128 (P.V, dS.K); 3 reduce only k >= m0 (P^T.dO, dS^T.Q).  tile_n 0 = auto.
 * batch > 1 with strideA (strideB) == 0: that operand is shared by every batch entry (grouped-query attention:
 * one K/V head serves several Q heads, models/int_llama_layer.py:138-139 repeat_kv).  accumulate (fp32 C only):
 * C += result (sums the K/V gradients over the Q heads of a group).
 * Replaces F.linear in QuantLinear.forward (quantize/int_linear.py:62) + its autograd dgrad/wgrad, and
 * torch.matmul/bmm in QuantMatMul.forward (quantize/int_matmul.py:41-43). */
int oq_gemm_bf16(const void* A, int a_mn, long lda, long strideA, const void* B, int b_mn, long ldb, long strideB,
                 void* C, int c_dtype, long ldc, long strideC, int M, int N, int K, int batch, const float* bias,
                 const void* residual, long ldr, long strideR, float alpha, int causal, int accumulate, int tile_n,
                 int max_ctas, oq_stream_t stream);

/* ---- K5/K6: fused causal attention of QuantMatMul (QK^T -> softmax -> PV in one kernel; recompute backward) --------
 * Replaces qkt_matmul / softmax / pv_matmul of QuantLlamaAttention.forward (models/int_llama_layer.py:141-163; OPT
 * models/int_opt_layer.py:124-197, Falcon models/int_falcon_layer.py:148-176) and their autograd on the already rotated +
 * fake-quantised operands.  qkv: [batch*T, ld] bf16 holding every head's q / k / v rows (`width` valid columns);
 * q of head h at column cq + (h / rep)*gq + (h % rep)*head_dim, k / v of K/V head g at ck / cv + g*gk (rep = heads /
 * kv_heads; LLaMA / OPT [q | k | v]: cq 0, gq rep*d, ck H, cv H+Hkv, gk d; Falcon per group [rep q | k | v]: gq = gk =
 * (rep+2)*d, ck rep*d, cv (rep+1)*d).  Scores are alpha*(q.k) rounded to fp16, softmax in fp32 with the causal mask,
 * P rounded to bf16 after normalisation.  O: [batch*T, heads*head_dim] bf16.  row_max / row_invl: [batch, heads, T] fp32
 * softmax statistics saved for the backward (nothing of size T x T is written).  Needs head_dim 64 or 128 and T %% 128 == 0
 * (oq_attn_supported); other shapes use oq_gemm_bf16 + oq_softmax_fwd/bwd.
 * Backward: dqkv (same column layout as qkv, row stride ldg) receives dq, dk, dv of every head (each element written
 * once; K/V heads shared by several query heads are summed inside the kernel); delta_ws: [batch, heads, T] fp32 scratch. */
int oq_attn_supported(int T, int head_dim);
int oq_attn_fwd(const void* qkv, long ld, long width, int batch, int T, int heads, int kv_heads, int head_dim, int cq, int gq,
                int ck, int cv, int gk, float alpha, void* O, long ldo, float* row_max, float* row_invl, oq_stream_t stream);
int oq_attn_bwd(const void* qkv, long ld, long width, int batch, int T, int heads, int kv_heads, int head_dim, int cq, int gq,
                int ck, int cv, int gk, float alpha, const void* O, long ldo, const void* dO, long lddo, const float* row_max,
                const float* row_invl, float* delta_ws, void* dqkv, long ldg, oq_stream_t stream);

/* ---- K1 as the north star words it: LWC fake-quant fused INTO the GEMM ------------------------------------------
 * oq_gemm_fq: C[M,N] (bf16) = A[M,K] (bf16) . fq(W)[N,K]^T (+ bias) (+ residual), where the fp32 master weight tile is
 * loaded by TMA, quantised / rounded / dequantised by transform warps in shared memory with the per-row (group_size 0) or
 * per-group scale / zero point [N, K/group_size], and consumed by tcgen05 from there (QuantLinear.forward with
 * use_weight_quant, quantize/int_linear.py:48-65 + quantize/quantizer.py:84-105).  The operand is bit-identical to
 * oq_weight_quant_fwd's bf16 output.  oq_lwc_qparams evaluates scale / zero point (quantizer.py:131-147) from the saved
 * group extrema and the LWC factors.  The engine uses the pre-pass + GEMM form instead; this entry point carries the
 * measured comparison (profiles/r2_k1_ab.txt).  K %% 64 == 0; group_size 0 or a multiple of 64. */
int oq_lwc_qparams(const float* xmin, const float* xmax, const float* up, const float* low, long n, int n_bits, int symmetric,
                   int recip_div, float* scale_out, float* zp_out, oq_stream_t stream);
int oq_gemm_fq(const void* A, long lda, const float* W, long ldw, int M, int N, int K, const float* scale, const float* zp,
               int group_size, int n_bits, void* C, long ldc, const float* bias, const void* residual, long ldr, int max_ctas,
               oq_stream_t stream);

/* ---- elementwise / reductions ----------------------------------------------
 * SwiGLU of QuantLlamaMLP.forward (models/int_llama_layer.py:44-45): gu = [gate | up] [rows, 2F]. */
int oq_swiglu_fwd(const void* gu, long ldgu, void* m, long ldm, long rows, int F, oq_stream_t stream);
int oq_swiglu_bwd(const void* gu, long ldgu, const void* dm, long lddm, void* dgu, long lddgu, long rows, int F,
                  oq_stream_t stream);
/* exact (erf) GELU of QuantFalconMLP.forward (nn.GELU(), models/int_falcon_layer.py:25,29) and its derivative. */
int oq_gelu_fwd(const void* x, void* y, long n, oq_stream_t stream);
int oq_gelu_bwd(const void* x, const void* dy, void* dx, long n, oq_stream_t stream);
/* ReLU of QuantOPTDecoderLayer.forward (models/int_opt_layer.py:318). */
int oq_relu_fwd(const void* x, void* y, long n, oq_stream_t stream);
int oq_relu_bwd(const void* x, const void* dy, void* dx, long n, oq_stream_t stream);
/* P = softmax_fp32(max(S + mask, finfo.min)) -> bf16 (models/int_llama_layer.py:156-160).  S [batch,Tq,Tk] in
 * fp32 (s_dtype 1) or fp16 (s_dtype 2). */
int oq_softmax_fwd(const void* S, int s_dtype, void* P, const float* mask, long mask_bstride, long batch, int Tq, int Tk,
                   int causal, oq_stream_t stream);
/* dS = P * (dP - rowsum(P*dP)) -> bf16 (dS may alias P); causal rows only read/write j <= i.  dP in bf16
 * (dp_dtype 0) or fp32 (1). */
int oq_softmax_bwd(const void* P, const void* dP, int dp_dtype, void* dS, long batch, int Tq, int Tk, int causal,
                   oq_stream_t stream);
/* loss_acc += mean((out-target)^2) [+ mean((out-target2)^2)]; dout = d loss / d out
 * (torch.nn.MSELoss at quantize/omniquant.py:142,220-222 and its backward). */
int oq_mse_fwd_bwd(const void* out, const void* target, const void* target2, long n, float inv_numel, float* loss_acc,
                   void* dout, oq_stream_t stream);
/* torch.optim.AdamW step on a flat fp32 segment (quantize/omniquant.py:207-208, utils.py:32-46);
 * g is multiplied by grad_scale first (1/world after the NCCL sum); *step_dev is the 1-based step. */
int oq_adamw(float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2, float eps,
             float weight_decay, float grad_scale, const int* step_dev, oq_stream_t stream);
int oq_step_inc(int* step_dev, oq_stream_t stream);
/* truncate_number on .data (models/transformation.py:5-20; models/int_llama_layer.py:281-284). */
int oq_truncate(float* p, long n, float threshold, oq_stream_t stream);
/* out[c] (+)= sum_r X[r,c]  (bias gradients feeding the LET shift/scale gradients). */
int oq_colsum(const void* X, long ld, long rows, int C, float* out, int accumulate, oq_stream_t stream);
int oq_transpose_bf16(const void* in, long ldi, void* out, long ldo, int R, int C, oq_stream_t stream);

/* ---- LET vector folds: the O(H) part of smooth_ln_fcs / smooth_fc_fc / smooth_q_k (models/transformation.py:
 * 28-33 norm weight/bias, :37-40,:56-59 bias + W@shift, :49-50 v bias, :66-69 q/k bias) and of their gradients,
 * one launch each.  Packed fp32 layouts (H hidden size, F1 rows of the first MLP weight group):
 *   let   [7H]    qkv_s | out_s | fc1_s | qkt_s | qkv_d | out_d | fc1_d          (parameters / gradients)
 *   mv    [4H+F1] W_q@d_qkv | W_k@d_qkv | W_v@d_qkv | W_o@d_out | W_fc1@d_fc1    (oq_weight_quant_fwd matvec_out)
 *   tmpv  [9H+F1] ln1.w' | ln1.b' | ln2.w' | ln2.b' | b_q' | b_k' | b_v' | b_o' | b_fc1' [F1] | b_q'*q_bias_scale
 *   acc   [10H]   dcol_qkv | dcol_out | dcol_fc1 | dsh_qkv | dsh_out | dsh_fc1 | dw1' | db1' | dw2' | db2'
 *   drow  [3H]    sum_k dW'*(W*col) of the q | k | v weights;   dbqkv [3H] column sums of dL/d(qkv)
 * part: 0 everything, 1 norm vectors + q/k/v biases, 2 o / fc1 biases (lets the two halves run on different streams).
 * q_bias_scale: head_dim^-0.5 for OPT (q is scaled before QuantMatMul, models/int_opt_layer.py:96), else 1. */
int oq_let_fold_vec(int H, int F1, const float* let_params, const float* ln1w, const float* ln1b, const float* ln2w,
                    const float* ln2b, const float* rawb_qkv, const float* rawb_o, const float* rawb_fc1, const float* mv,
                    float q_bias_scale, float* tmpv, int part, oq_stream_t stream);
int oq_let_prep_bwd(int H, const float* let_params, const float* dbqkv, float q_bias_scale, float* e, oq_stream_t stream);
int oq_let_grad_vec(int H, int F1, const float* let_params, const float* tmpv, const float* acc, const float* drow,
                    const float* dbqkv, float q_bias_scale, float* let_grads, oq_stream_t stream);

/* ---- activation statistics for the LET initialisation (next row f3 of SURVEY.md section 8)
 * Replaces generate_act_scale_shift.py:25-58 (get_act_scales: per-channel running max of |x| over samples) and :60-94
 * (get_act_shifts: (max+min)/2 of the first sample, then 0.99*prev + 0.01*(max+min)/2 per sample) for the input of
 * one linear layer.  oq_act_stats reduces ONE sample x[rows, C] (bf16, row stride ld) to absmax[C] and mid[C]
 * (either may be NULL); workspace of *bytes_out bytes as reported by
 * oq_act_stats_workspace(rows, C, &bytes).  oq_act_stats_fold folds the vectors of
 * n samples (absmax[n, C], mid[n, C], sample order) into scale_io[C] / shift_io[C]; first != 0 starts the
 * recursion from sample 0 (the reference's `name not in act_scales` branch), otherwise the running values continue. */
int oq_act_stats_workspace(int rows, int C, long* bytes_out);
int oq_act_stats(const void* x, int rows, int C, long ld, float* absmax, float* mid, float* workspace, long workspace_bytes,
                 oq_stream_t stream);
int oq_act_stats_fold(const float* absmax, const float* mid, int n, int C, int first, float* scale_io, float* shift_io,
                      oq_stream_t stream);

/* ---- real-quant export (next row f1 of SURVEY.md section 8)
 * Replaces `qlinear_cuda.QuantLinear(wbits, group_size, in, out, bias).pack(module, scales, zeros)` at
 * quantize/omniquant.py:255-265 (AutoGPTQ wire format; see oracle/gptq_pack.py for the restated algorithm and what is
 * unpinned).  W[N, K] is the fake-quantised fp32 weight, scales / zeros [N, G] (G = K / group_size, group_size 0 =
 * per channel) as registered by quantizer.py:149-153.  Outputs: qweight int32 [K/32*bits, N], qzeros int32
 * [G, N/32*bits] (zero - 1, fields masked to `bits` bits), scales_f16 [G, N], g_idx int32 [K].  bits in {2,3,4,8};
 * N and K multiples of 32. */
int oq_gptq_pack(const float* W, const float* scales, const float* zeros, int N, int K, int group_size, int n_bits,
                 int32_t* qweight, int32_t* qzeros, void* scales_f16, int32_t* g_idx, oq_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OMNIQUANT_B200_H_ */
