"""Thin tensor-level wrappers over the C ABI (include/omniquant_b200.h).

Every function takes torch CUDA tensors (device memory + the current stream
are the only things torch contributes), checks dtypes / layouts, and enqueues
exactly one library call.  No computation happens in Python.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from ._lib import lib, OqError

BF16, F32 = torch.bfloat16, torch.float32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk(t: Optional[torch.Tensor], dtype, name: str):
    if t is None:
        return
    if not t.is_cuda:
        raise OqError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise OqError(f"{name} must be {dtype}, got {t.dtype}")


def _rowmajor(t: torch.Tensor, name: str):
    if t.dim() != 2 or t.stride(1) != 1:
        raise OqError(f"{name} must be 2-D with unit inner stride, got shape {tuple(t.shape)} strides {t.stride()}")
    return t.stride(0)


# ------------------------------------------------------------------ weight path
class S8Operand:
    """Exact-integer operand of gemm_s8 as emitted by a quantiser kernel: signed codes [rows, C] int8 (code - z0) and the
    per-row vectors scale, t = scale * sum_k (code - z0), beta = (z0 - zero_point) * scale (views of one [3, rows] buffer)."""

    def __init__(self, rows, C, device):
        self.codes = torch.empty(rows, C, device=device, dtype=torch.int8)
        self.meta = torch.empty(3, rows, device=device, dtype=F32)

    @property
    def scale(self):
        return self.meta[0]

    @property
    def t(self):
        return self.meta[1]

    @property
    def beta(self):
        return self.meta[2]

    def rows(self, r0, r1):
        """The operand restricted to rows [r0, r1) (shares storage)."""
        v = object.__new__(S8Operand)
        v.codes, v.meta = self.codes[r0:r1], self.meta[:, r0:r1]
        return v

    def clone(self):
        v = object.__new__(S8Operand)
        v.codes, v.meta = self.codes.clone(), self.meta.clone()
        return v


def _s8_args(s8, rows, C):
    if s8 is None:
        return None, 0, None
    _chk(s8.codes, torch.int8, "s8.codes")
    _chk(s8.meta, F32, "s8.meta")
    if tuple(s8.codes.shape) != (rows, C) or s8.codes.stride(1) != 1 or tuple(s8.meta.shape) != (3, rows) or not s8.meta.is_contiguous():
        raise OqError(f"s8 operand must hold codes [{rows}, {C}] and a contiguous meta [3, {rows}]")
    return s8.codes.data_ptr(), s8.codes.stride(0), s8.meta.data_ptr()


def _segments(segments, N):
    """(n_seg, seg_end, seg_row, seg_op) ctypes arguments of a row segment table [(end_row, row_scale | None, op), ...]."""
    import ctypes
    if not segments:
        return 0, None, None, None
    if len(segments) > 3 or segments[-1][0] != N:
        raise OqError("row segments: at most 3, the last one must end at N")
    for _, t, _ in segments:
        _chk(t, F32, "segment row scale")
    ends = (ctypes.c_int * len(segments))(*[int(e) for e, _, _ in segments])
    rows = (ctypes.c_void_p * len(segments))(*[(t.data_ptr() if t is not None else None) for _, t, _ in segments])
    opsv = (ctypes.c_int * len(segments))(*[int(o) for _, _, o in segments])
    return len(segments), ends, rows, opsv


def weight_quant_fwd(W, *, n_bits, symmetric=False, group_size=0, up=None, low=None, col_scale=None, row_scale=None,
                     row_op=0, shift=None, matvec_out=None, recip_div=True, wq_bf16=None, wq_f32=None, codes=None,
                     scale_out=None, zp_out=None, xmin_out=None, xmax_out=None, s8=None, segments=None):
    _chk(W, F32, "W")
    for n, t in (("up", up), ("low", low), ("col_scale", col_scale), ("row_scale", row_scale), ("shift", shift),
                 ("matvec_out", matvec_out), ("wq_f32", wq_f32), ("scale_out", scale_out), ("zp_out", zp_out),
                 ("xmin_out", xmin_out), ("xmax_out", xmax_out)):
        _chk(t, F32, n)
    _chk(wq_bf16, BF16, "wq_bf16")
    _chk(codes, torch.uint8, "codes")
    if not W.is_contiguous():
        raise OqError("W must be contiguous [N, K]")
    N, K = W.shape
    ld = _rowmajor(wq_bf16, "wq_bf16") if wq_bf16 is not None else 0
    c8 = t8 = b8 = None
    if s8 is not None:      # the weight's scale goes to s8.meta[0] (== scale_out when the caller did not ask for its own)
        if tuple(s8.codes.shape) != (N, K) or not s8.codes.is_contiguous():
            raise OqError("weight_quant_fwd: s8 operand must hold contiguous codes [N, K]")
        if scale_out is not None and scale_out.data_ptr() != s8.scale.data_ptr():
            raise OqError("weight_quant_fwd: with an s8 operand scale_out must be s8.scale (or None)")
        scale_out, c8, t8, b8 = s8.scale, s8.codes.data_ptr(), s8.t.data_ptr(), s8.beta.data_ptr()
    lib().call("oq_weight_quant_fwd", _p(W), N, K, _p(col_scale), _p(row_scale), row_op, _p(shift), _p(matvec_out),
               n_bits, int(symmetric), int(group_size or 0), _p(up), _p(low), int(recip_div), _p(wq_bf16), ld,
               _p(wq_f32), _p(codes), _p(scale_out), _p(zp_out), _p(xmin_out), _p(xmax_out), c8, t8, b8,
               *_segments(segments, N), _stream())


def weight_quant_bwd(W, grad_wq, xmin, xmax, *, n_bits, symmetric=False, group_size=0, up=None, low=None,
                     col_scale=None, row_scale=None, row_op=0, recip_div=True, e_row=None, d_up=None, d_low=None,
                     dcol_acc=None, drow=None, dshift_acc=None, dW_full=None, scale=None, zp=None, segments=None):
    _chk(W, F32, "W")
    if grad_wq.dtype not in (BF16, F32):
        raise OqError("grad_wq must be bf16 or fp32")
    N, K = W.shape
    ldg = _rowmajor(grad_wq, "grad_wq")
    lib().call("oq_weight_quant_bwd", _p(W), N, K, _p(col_scale), _p(row_scale), row_op, n_bits, int(symmetric),
               int(group_size or 0), _p(up), _p(low), int(recip_div), _p(xmin), _p(xmax), _p(scale), _p(zp), _p(grad_wq),
               int(grad_wq.dtype == F32), ldg, _p(e_row), _p(d_up), _p(d_low), _p(dcol_acc), _p(drow), _p(dshift_acc),
               _p(dW_full), *_segments(segments, N), _stream())


# ------------------------------------------------------------------ activations
def _rows2(x: torch.Tensor, name: str):
    """(rows, C, rows_inner, ld_outer, ld_inner) for a 2-D [rows, C] or 3-D [outer, inner, C] bf16 view."""
    _chk(x, BF16, name)
    if x.stride(-1) != 1:
        raise OqError(f"{name}: last dim must be contiguous")
    if x.dim() == 2:
        return x.shape[0], x.shape[1], 1, x.stride(0), 0
    if x.dim() == 3:
        return x.shape[0] * x.shape[1], x.shape[2], x.shape[1], x.stride(0), x.stride(1)
    raise OqError(f"{name}: expected 2-D or 3-D view, got {tuple(x.shape)}")


def act_quant_fwd(x, y, *, n_bits, symmetric=False, recip_div=True, cos=None, sin=None, seq_len=0, rope_inner=0,
                  scale_out=None, zp_out=None, s8=None):
    rows, C, ri, lo, li = _rows2(x, "x")
    rows2, C2, ri2, lo2, li2 = _rows2(y, "y")
    if (rows, C, ri) != (rows2, C2, ri2):
        raise OqError("act_quant_fwd: x / y shape mismatch")
    c8, ld8, m8 = _s8_args(s8, rows, C)
    lib().call("oq_act_quant_fwd", _p(x), rows, C, ri, lo, li, n_bits, int(symmetric), int(recip_div), _p(cos), _p(sin),
               seq_len, rope_inner, _p(y), lo2, li2, _p(scale_out), _p(zp_out), c8, ld8, m8, _stream())


def act_quant_bwd(x, gy, gx, *, n_bits, symmetric=False, recip_div=True, cos=None, sin=None, seq_len=0, rope_inner=0):
    rows, C, ri, lo, li = _rows2(x, "x")
    r2, c2, ri2, glo, gli = _rows2(gy, "gy")
    r3, c3, ri3, xlo, xli = _rows2(gx, "gx")
    if (rows, C, ri) != (r2, c2, ri2) or (rows, C, ri) != (r3, c3, ri3):
        raise OqError("act_quant_bwd: shape mismatch")
    lib().call("oq_act_quant_bwd", _p(x), rows, C, ri, lo, li, n_bits, int(symmetric), int(recip_div), _p(cos), _p(sin),
               seq_len, rope_inner, _p(gy), glo, gli, _p(gx), xlo, xli, _stream())


def swiglu_act_quant_fwd(gu, mq, *, n_bits, symmetric=False, recip_div=True, s8=None):
    """mq = fake_quant(silu(gate) * up) for gu = [gate | up]; the product itself is not stored."""
    _chk(gu, BF16, "gu"); _chk(mq, BF16, "mq")
    rows, F = mq.shape
    if gu.shape != (rows, 2 * F):
        raise OqError("swiglu_act_quant_fwd: gu must be [rows, 2F]")
    c8, ld8, m8 = _s8_args(s8, rows, F)
    lib().call("oq_swiglu_act_quant_fwd", _p(gu), _rowmajor(gu, "gu"), rows, F, n_bits, int(symmetric), int(recip_div),
               _p(mq), _rowmajor(mq, "mq"), c8, ld8, m8, _stream())


def swiglu_act_quant_bwd(gu, gy, dgu, *, n_bits, symmetric=False, recip_div=True):
    _chk(gu, BF16, "gu"); _chk(gy, BF16, "gy"); _chk(dgu, BF16, "dgu")
    rows, F = gy.shape
    if gu.shape != (rows, 2 * F) or dgu.shape != (rows, 2 * F):
        raise OqError("swiglu_act_quant_bwd: gu / dgu must be [rows, 2F]")
    lib().call("oq_swiglu_act_quant_bwd", _p(gu), _rowmajor(gu, "gu"), rows, F, n_bits, int(symmetric), int(recip_div),
               _p(gy), _rowmajor(gy, "gy"), _p(dgu), _rowmajor(dgu, "dgu"), _stream())


# ------------------------------------------------------------------ norms
def norm_fwd(x, w, b, eps, kind, y, mean_out=None, rstd_out=None):
    _chk(x, BF16, "x"); _chk(y, BF16, "y"); _chk(w, F32, "w"); _chk(b, F32, "b")
    lib().call("oq_norm_fwd", _p(x), x.shape[0], x.shape[1], _rowmajor(x, "x"), _p(w), _p(b), float(eps), kind, _p(y),
               _rowmajor(y, "y"), _p(mean_out), _p(rstd_out), _stream())


def norm_act_quant_fwd(x, w, b, eps, kind, y, yq, *, n_bits, symmetric=False, recip_div=True, mean_out=None, rstd_out=None,
                       s8=None):
    """y = norm(x) (optional) and yq = fake_quant(y) per token, one pass."""
    _chk(x, BF16, "x"); _chk(y, BF16, "y"); _chk(yq, BF16, "yq"); _chk(w, F32, "w"); _chk(b, F32, "b")
    lib().call("oq_norm_act_quant_fwd", _p(x), x.shape[0], x.shape[1], _rowmajor(x, "x"), _p(w), _p(b), float(eps), kind,
               _p(y), _rowmajor(y, "y") if y is not None else 0, _p(mean_out), _p(rstd_out), n_bits, int(symmetric),
               int(recip_div), _p(yq), _rowmajor(yq, "yq"), *_s8_args(s8, x.shape[0], x.shape[1]), _stream())


def norm_bwd(x, w, kind, mean, rstd, dy, dx=None, dx_add=None, dw_acc=None, db_acc=None):
    _chk(x, BF16, "x"); _chk(dy, BF16, "dy"); _chk(dx, BF16, "dx"); _chk(dx_add, BF16, "dx_add")
    lib().call("oq_norm_bwd", _p(x), x.shape[0], x.shape[1], _rowmajor(x, "x"), _p(w), kind, _p(mean), _p(rstd), _p(dy),
               _rowmajor(dy, "dy"), _p(dx), _rowmajor(dx, "dx") if dx is not None else 0, _p(dx_add),
               _rowmajor(dx_add, "dx_add") if dx_add is not None else 0, _p(dw_acc), _p(db_acc), _stream())


# ------------------------------------------------------------------ GEMM
_CDT = {BF16: 0, F32: 1, torch.float16: 2}
GEMM_TIMER = None   # bench.py sets this to a list to collect (event0, event1, flops, kind: lin | attn | s8) per GEMM launch


def gemm_s8(A, B, C, row_scale, col_scale, *, corr=None, alpha=1.0, bias=None, residual=None):
    """C = alpha * row_scale[:, None] * col_scale[None, :] * (A.int() @ B.int().T) (+ zero-point corrections) (+ bias)
    (+ residual) for int8 code matrices A [M, K], B [N, K] (exact integer sum on the tensor cores);
    corr = (row_t, row_beta, col_t, col_beta) or None."""
    if A.dtype != torch.int8 or B.dtype != torch.int8 or A.dim() != 2 or B.dim() != 2 or A.shape[1] != B.shape[1]:
        raise OqError("gemm_s8: A [M, K] and B [N, K] must be int8")
    if A.stride(1) != 1 or B.stride(1) != 1 or C.stride(1) != 1:
        raise OqError("gemm_s8: last dims must be contiguous")
    M, K = A.shape
    N = B.shape[0]
    rt, rb, ct, cb = corr if corr is not None else (None, None, None, None)
    if GEMM_TIMER is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    lib().call("oq_gemm_s8", _p(A), A.stride(0), _p(B), B.stride(0), _p(C), _CDT[C.dtype], C.stride(0), M, N, K,
               _p(row_scale), _p(col_scale), _p(rt), _p(rb), _p(ct), _p(cb), float(alpha), _p(bias), _p(residual),
               residual.stride(0) if residual is not None else 0, _stream())
    if GEMM_TIMER is not None:
        e1.record()
        GEMM_TIMER.append((e0, e1, 2.0 * M * N * K, "s8"))


def linear_s8_fwd(xs: "S8Operand", ws: "S8Operand", out, bias=None, residual=None, alpha=1.0):
    """out[M,N] = alpha * x_q[M,K] @ w_q[N,K]^T + bias + residual from the integer codes of both fake-quantised operands
    (QuantLinear.forward with use_act_quant and a quantised weight, quantize/int_linear.py:48-65)."""
    gemm_s8(xs.codes, ws.codes, out, xs.scale, ws.scale, corr=(xs.t, xs.beta, ws.t, ws.beta), alpha=alpha, bias=bias,
            residual=residual)


def launch_count() -> int:
    return lib().launches


def gemm(A, B, C, *, M, N, K, a_mn=False, b_mn=False, lda=None, ldb=None, ldc=None, batch=1, strideA=0, strideB=0,
         strideC=0, bias=None, residual=None, ldr=0, strideR=0, alpha=1.0, causal=0, tile_n=0, max_ctas=0,
         accumulate=False):
    """C[b] = alpha * A[b] @ B[b]^T (+bias) (+residual); raw strided form (see the header)."""
    _chk(A, BF16, "A"); _chk(B, BF16, "B"); _chk(bias, F32, "bias"); _chk(residual, BF16, "residual")
    if C.dtype not in _CDT:
        raise OqError("C must be bf16, fp32 or fp16")
    if GEMM_TIMER is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    lib().call("oq_gemm_bf16", _p(A), int(a_mn), lda, strideA, _p(B), int(b_mn), ldb, strideB, _p(C),
               _CDT[C.dtype], ldc, strideC, M, N, K, batch, _p(bias), _p(residual), ldr, strideR, float(alpha),
               causal, int(accumulate), tile_n, max_ctas, _stream())
    if GEMM_TIMER is not None:
        e1.record()
        GEMM_TIMER.append((e0, e1, 2.0 * M * N * K * batch * (0.5 if causal else 1.0), "attn" if batch > 1 else "lin"))


def linear_fwd(x, wq, out, bias=None, residual=None):
    """out[M,N] = x[M,K] @ wq[N,K]^T + bias + residual   (QuantLinear forward GEMM)."""
    M, K = x.shape
    N = wq.shape[0]
    gemm(x, wq, out, M=M, N=N, K=K, lda=_rowmajor(x, "x"), ldb=_rowmajor(wq, "wq"), ldc=_rowmajor(out, "out"), bias=bias,
         residual=residual, ldr=_rowmajor(residual, "residual") if residual is not None else 0)


def linear_dgrad(dy, wq, dx, residual=None):
    """dx[M,K] = dy[M,N] @ wq[N,K]   (B operand read N-major, no transpose copy)."""
    M, N = dy.shape
    K = wq.shape[1]
    gemm(dy, wq, dx, M=M, N=K, K=N, lda=_rowmajor(dy, "dy"), b_mn=True, ldb=_rowmajor(wq, "wq"), ldc=_rowmajor(dx, "dx"),
         residual=residual, ldr=_rowmajor(residual, "residual") if residual is not None else 0)


def linear_wgrad(dy, x, dw):
    """dw[N,K] = dy[M,N]^T @ x[M,K]   (both operands read MN-major)."""
    M, N = dy.shape
    K = x.shape[1]
    gemm(dy, x, dw, M=N, N=K, K=M, a_mn=True, lda=_rowmajor(dy, "dy"), b_mn=True, ldb=_rowmajor(x, "x"),
         ldc=_rowmajor(dw, "dw"))


def lwc_qparams(xmin, xmax, up, low, *, n_bits, symmetric=False, recip_div=True, scale_out=None, zp_out=None):
    """scale / zero point of every (row, group) from the saved extrema and the LWC factors (quantizer.py:131-147)."""
    for n, t in (("xmin", xmin), ("xmax", xmax), ("up", up), ("low", low)):
        _chk(t, F32, n)
    scale_out = torch.empty_like(xmin) if scale_out is None else scale_out
    zp_out = torch.empty_like(xmin) if zp_out is None else zp_out
    lib().call("oq_lwc_qparams", _p(xmin), _p(xmax), _p(up), _p(low), xmin.numel(), n_bits, int(symmetric), int(recip_div),
               _p(scale_out), _p(zp_out), _stream())
    return scale_out, zp_out


def gemm_fq(x, W, scale, zp, out, *, n_bits, group_size=0, bias=None, residual=None, max_ctas=0):
    """out[M,N] = x[M,K] @ fake_quant(W)[N,K]^T with the weight quantiser fused into the GEMM's operand path (K1)."""
    _chk(x, BF16, "x"); _chk(W, F32, "W"); _chk(out, BF16, "out"); _chk(scale, F32, "scale"); _chk(zp, F32, "zp")
    _chk(bias, F32, "bias"); _chk(residual, BF16, "residual")
    M, K = x.shape
    N = W.shape[0]
    lib().call("oq_gemm_fq", _p(x), _rowmajor(x, "x"), _p(W), _rowmajor(W, "W"), M, N, K, _p(scale), _p(zp), int(group_size or 0),
               n_bits, _p(out), _rowmajor(out, "out"), _p(bias), _p(residual),
               _rowmajor(residual, "residual") if residual is not None else 0, max_ctas, _stream())


# ------------------------------------------------------------------ fused attention
def attn_supported(T: int, head_dim: int) -> bool:
    return bool(lib().fns["oq_attn_supported"](int(T), int(head_dim)))


def _attn_common(qkv, layout):
    _chk(qkv, BF16, "qkv")
    if qkv.dim() != 2 or qkv.stride(1) != 1:
        raise OqError("attention: qkv must be a 2-D [batch*T, width] view with contiguous columns")
    return tuple(int(layout[k]) for k in ("cq", "gq", "ck", "cv", "gk"))


def attn_fwd(qkv, O, row_max, row_invl, *, batch, T, heads, kv_heads, head_dim, layout, alpha):
    """O = softmax_causal(alpha * q k^T) v for every (sample, head) of the fused operand matrix, one launch."""
    cq, gq, ck, cv, gk = _attn_common(qkv, layout)
    _chk(O, BF16, "O"); _chk(row_max, F32, "row_max"); _chk(row_invl, F32, "row_invl")
    lib().call("oq_attn_fwd", _p(qkv), qkv.stride(0), qkv.shape[1], batch, T, heads, kv_heads, head_dim, cq, gq, ck, cv, gk,
               float(alpha), _p(O), _rowmajor(O, "O"), _p(row_max), _p(row_invl), _stream())


def attn_bwd(qkv, O, dO, row_max, row_invl, delta_ws, dqkv, *, batch, T, heads, kv_heads, head_dim, layout, alpha):
    """dq | dk | dv into dqkv (same column layout as qkv); recomputes the probabilities from the saved row statistics."""
    cq, gq, ck, cv, gk = _attn_common(qkv, layout)
    for n, t in (("O", O), ("dO", dO), ("dqkv", dqkv)):
        _chk(t, BF16, n)
    for n, t in (("row_max", row_max), ("row_invl", row_invl), ("delta_ws", delta_ws)):
        _chk(t, F32, n)
    lib().call("oq_attn_bwd", _p(qkv), qkv.stride(0), qkv.shape[1], batch, T, heads, kv_heads, head_dim, cq, gq, ck, cv, gk,
               float(alpha), _p(O), _rowmajor(O, "O"), _p(dO), _rowmajor(dO, "dO"), _p(row_max), _p(row_invl), _p(delta_ws),
               _p(dqkv), _rowmajor(dqkv, "dqkv"), _stream())


# ------------------------------------------------------------------ elementwise
def swiglu_fwd(gu, m):
    lib().call("oq_swiglu_fwd", _p(gu), _rowmajor(gu, "gu"), _p(m), _rowmajor(m, "m"), m.shape[0], m.shape[1], _stream())


def swiglu_bwd(gu, dm, dgu):
    lib().call("oq_swiglu_bwd", _p(gu), _rowmajor(gu, "gu"), _p(dm), _rowmajor(dm, "dm"), _p(dgu), _rowmajor(dgu, "dgu"),
               dm.shape[0], dm.shape[1], _stream())


def gelu_fwd(x, y):
    lib().call("oq_gelu_fwd", _p(x), _p(y), x.numel(), _stream())


def gelu_bwd(x, dy, dx):
    lib().call("oq_gelu_bwd", _p(x), _p(dy), _p(dx), x.numel(), _stream())


def relu_fwd(x, y):
    lib().call("oq_relu_fwd", _p(x), _p(y), x.numel(), _stream())


def relu_bwd(x, dy, dx):
    lib().call("oq_relu_bwd", _p(x), _p(dy), _p(dx), x.numel(), _stream())


def softmax_fwd(S, P, mask=None, causal=False):
    """S fp32 or fp16 [B,Tq,Tk] contiguous -> P bf16."""
    _chk(P, BF16, "P"); _chk(mask, F32, "mask")
    if S.dtype not in (F32, torch.float16):
        raise OqError("S must be fp32 or fp16")
    B, Tq, Tk = S.shape
    mbs = 0
    if mask is not None:
        mask = mask.reshape(-1, Tq, Tk)
        mbs = 0 if mask.shape[0] == 1 else Tq * Tk
    lib().call("oq_softmax_fwd", _p(S), 1 if S.dtype == F32 else 2, _p(P), _p(mask), mbs, B, Tq, Tk, int(causal), _stream())


def softmax_bwd(P, dP, dS, causal=False):
    _chk(P, BF16, "P"); _chk(dS, BF16, "dS")
    if dP.dtype not in (F32, BF16):
        raise OqError("dP must be fp32 or bf16")
    B, Tq, Tk = P.shape
    lib().call("oq_softmax_bwd", _p(P), _p(dP), int(dP.dtype == F32), _p(dS), B, Tq, Tk, int(causal), _stream())


def mse_fwd_bwd(out, target, loss_acc, dout=None, target2=None):
    _chk(out, BF16, "out"); _chk(target, BF16, "target"); _chk(target2, BF16, "target2"); _chk(loss_acc, F32, "loss_acc")
    n = out.numel()
    lib().call("oq_mse_fwd_bwd", _p(out), _p(target), _p(target2), n, 1.0 / n, _p(loss_acc), _p(dout), _stream())


def adamw(p, g, m, v, step_dev, *, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, grad_scale=1.0):
    for n, t in (("p", p), ("g", g), ("m", m), ("v", v)):
        _chk(t, F32, n)
    lib().call("oq_adamw", _p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps, weight_decay, grad_scale,
               _p(step_dev), _stream())


def step_inc(step_dev):
    lib().call("oq_step_inc", _p(step_dev), _stream())


def truncate_(p, threshold=1e-2):
    _chk(p, F32, "p")
    lib().call("oq_truncate", _p(p), p.numel(), threshold, _stream())


def colsum(X, out, accumulate=False):
    _chk(X, BF16, "X"); _chk(out, F32, "out")
    lib().call("oq_colsum", _p(X), _rowmajor(X, "X"), X.shape[0], X.shape[1], _p(out), int(accumulate), _stream())


def transpose(x, out):
    lib().call("oq_transpose_bf16", _p(x), _rowmajor(x, "x"), _p(out), _rowmajor(out, "out"), x.shape[0], x.shape[1],
               _stream())


# ------------------------------------------------------------------ LET vector folds
def let_fold_vec(H, F1, let_params, ln1w, ln1b, ln2w, ln2b, rawb_qkv, rawb_o, rawb_fc1, mv, q_bias_scale, tmpv, part=0):
    lib().call("oq_let_fold_vec", H, F1, _p(let_params), _p(ln1w), _p(ln1b), _p(ln2w), _p(ln2b), _p(rawb_qkv), _p(rawb_o),
               _p(rawb_fc1), _p(mv), float(q_bias_scale), _p(tmpv), int(part), _stream())


def let_prep_bwd(H, let_params, dbqkv, q_bias_scale, e):
    lib().call("oq_let_prep_bwd", H, _p(let_params), _p(dbqkv), float(q_bias_scale), _p(e), _stream())


def let_grad_vec(H, F1, let_params, tmpv, acc, drow, dbqkv, q_bias_scale, let_grads):
    lib().call("oq_let_grad_vec", H, F1, _p(let_params), _p(tmpv), _p(acc), _p(drow), _p(dbqkv), float(q_bias_scale),
               _p(let_grads), _stream())


# ------------------------------------------------------------------ activation statistics (LET initialisation)
_STATS_WS = {}


def act_stats(x: torch.Tensor, absmax: torch.Tensor = None, mid: torch.Tensor = None):
    """Per-channel max|x| and (max+min)/2 over the rows of one sample x [rows, C] (bf16)
    (generate_act_scale_shift.py:30-33 and :64-72)."""
    import ctypes
    _chk(x, BF16, "x")
    _chk(absmax, F32, "absmax")
    _chk(mid, F32, "mid")
    if x.dim() != 2 or x.stride(1) != 1:
        raise OqError("act_stats: x must be a 2-D view with contiguous channels")
    rows, C = x.shape
    key = (rows, C, x.device)
    ws = _STATS_WS.get(key)
    if ws is None:
        nbytes = ctypes.c_long(0)
        lib().call("oq_act_stats_workspace", rows, C, ctypes.addressof(nbytes))
        ws = _STATS_WS[key] = torch.empty(nbytes.value // 4, device=x.device, dtype=F32)
    lib().call("oq_act_stats", _p(x), rows, C, x.stride(0), _p(absmax), _p(mid), _p(ws), ws.numel() * 4, _stream())


def act_stats_fold(absmax, mid, scale_io=None, shift_io=None, first=True):
    """Fold per-sample vectors (absmax [n, C], mid [n, C], sample order) into the running max / EMA of the reference."""
    t = absmax if absmax is not None else mid
    n, C = t.shape
    for name, v in (("absmax", absmax), ("mid", mid), ("scale_io", scale_io), ("shift_io", shift_io)):
        _chk(v, F32, name)
        if v is not None and not v.is_contiguous():
            raise OqError(f"act_stats_fold: {name} must be contiguous")
    lib().call("oq_act_stats_fold", _p(absmax), _p(mid), n, C, int(first), _p(scale_io), _p(shift_io), _stream())


# ------------------------------------------------------------------ real-quant export (GPTQ wire format)
def gptq_pack(W: torch.Tensor, scales: torch.Tensor, zeros: torch.Tensor, *, n_bits: int, group_size: int = 0):
    """(qweight int32 [K/32*bits, N], qzeros int32 [G, N/32*bits], scales fp16 [G, N], g_idx int32 [K]) of a
    fake-quantised weight W [N, K] with its per-group scales / zero points [N*G(,1)] (quantize/omniquant.py:255-265)."""
    _chk(W, F32, "W")
    _chk(scales, F32, "scales")
    _chk(zeros, F32, "zeros")
    if not (W.is_contiguous() and scales.is_contiguous() and zeros.is_contiguous()):
        raise OqError("gptq_pack: contiguous tensors required")
    N, K = W.shape
    group = group_size or K
    if K % group or K % 32 or N % 32:
        raise OqError(f"gptq_pack: N={N}, K={K} must be multiples of 32 and K a multiple of the group size {group}")
    G = K // group
    if scales.numel() != N * G or zeros.numel() != N * G:
        raise OqError(f"gptq_pack: scales / zeros must hold {N * G} values")
    dev = W.device
    qweight = torch.empty(K // 32 * n_bits, N, device=dev, dtype=torch.int32)
    qzeros = torch.empty(G, N // 32 * n_bits, device=dev, dtype=torch.int32)
    sc = torch.empty(G, N, device=dev, dtype=torch.float16)
    g_idx = torch.empty(K, device=dev, dtype=torch.int32)
    lib().call("oq_gptq_pack", _p(W), _p(scales), _p(zeros), N, K, int(group_size or 0), n_bits, _p(qweight), _p(qzeros),
               _p(sc), _p(g_idx), _stream())
    return qweight, qzeros, sc, g_idx
