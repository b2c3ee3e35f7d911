"""LET folding with the reference's function names and argument meaning (reference
models/transformation.py:5-114).

The six public entry points (``smooth_{ln_fcs,fc_fc,q_k}_{temporary,inplace}``) keep the
reference's side effects on the module attributes -- ``temp_weight`` / ``temp_bias`` /
``use_temporary_parameter`` for the differentiable per-step variant, in-place buffer updates (and
newly registered ``bias`` buffers) for the once-per-block variant -- but share one implementation
of the algebra.  Inside the calibration step the same algebra is fused into the weight fake-quant
kernel (csrc/weight_quant.cu: column scale, row scale / divide, W@shift) and never materialises
the folded fp32 weights.
"""
import torch

from .. import ops

__all__ = ["truncate_number", "smooth_ln_fcs_temporary", "smooth_fc_fc_temporary", "smooth_q_k_temporary",
           "smooth_ln_fcs_inplace", "smooth_fc_fc_inplace", "smooth_q_k_inplace", "TruncateFunction"]


class TruncateFunction(torch.autograd.Function):
    """|x| < thr -> sign(x) * thr with identity backward (transformation.py:5-16)."""

    @staticmethod
    def forward(ctx, input, threshold):
        clipped = input.detach().clone()
        if clipped.is_cuda and clipped.dtype == torch.float32:
            ops.truncate_(clipped.view(-1), threshold)
        else:
            tiny = clipped.abs() < threshold
            clipped[tiny] = clipped[tiny].sign() * threshold
        return clipped

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.clone(), None


def truncate_number(number, threshold=1e-2):
    return TruncateFunction.apply(number, threshold)


# ---------------------------------------------------------------------------- shared algebra
class _FoldFn(torch.autograd.Function):
    """(W * col[k]) {none | * row[n] | / row[n]} and W @ shift in ONE pass of the fold-only mode of the weight kernel
    (oq_weight_quant_fwd with n_bits = 16: identity quantiser), differentiable w.r.t. col / row / shift through the
    kernel's backward (oq_weight_quant_bwd: column, row and shift reductions).  CUDA fp32 only -- no CPU path."""

    @staticmethod
    def forward(ctx, W, col, row, row_op, shift):
        if not (W.is_cuda and W.dtype == torch.float32):
            raise ops.OqError("LET folds run on CUDA fp32 weights only (there is no CPU path)")
        f = lambda t: None if t is None else t.detach().to(W.device, torch.float32).contiguous().view(-1)
        Wc = W.detach().contiguous()
        colc, rowc, shc = f(col), f(row), f(shift)
        out = torch.empty_like(Wc)
        mv = torch.empty(Wc.shape[0], device=W.device, dtype=torch.float32) if shc is not None else None
        ops.weight_quant_fwd(Wc, n_bits=16, col_scale=colc, row_scale=rowc, row_op=row_op if rowc is not None else 0,
                             shift=shc, matvec_out=mv, wq_f32=out)
        ctx.save_for_backward(Wc, *[t if t is not None else torch.empty(0, device=W.device) for t in (colc, rowc, shc)])
        ctx.have = (colc is not None, rowc is not None, shc is not None)
        ctx.row_op = row_op
        ctx.shapes = tuple(None if t is None else t.shape for t in (col, row, shift))
        return out, mv

    @staticmethod
    def backward(ctx, gW, gmv):
        Wc, colc, rowc, shc = ctx.saved_tensors
        hc, hr, hs = ctx.have
        colc, rowc, shc = (colc if hc else None), (rowc if hr else None), (shc if hs else None)
        N, K = Wc.shape
        dev = Wc.device
        z = lambda n: torch.zeros(n, device=dev, dtype=torch.float32)
        gW = z(N * K).view(N, K) if gW is None else gW.to(torch.float32).contiguous()
        e = None
        if hs:
            e = z(N) if gmv is None else gmv.to(torch.float32).contiguous()
        dcol, drow, dsh = (z(K) if hc else None), (z(N) if hr else None), (z(K) if hs else None)
        need_w = ctx.needs_input_grad[0]
        dfull = torch.empty_like(Wc) if need_w else None
        dummy = z(N)
        ops.weight_quant_bwd(Wc, gW, dummy, dummy, n_bits=16, col_scale=colc, row_scale=rowc,
                             row_op=ctx.row_op if hr else 0, e_row=e, dcol_acc=dcol, drow=drow, dshift_acc=dsh, dW_full=dfull)
        if hr and ctx.row_op == 2:          # W' = W1 / r  ->  d/dr = -sum_k dW' * W1 / r^2
            drow = -drow / (rowc * rowc)
        dWin = None
        if need_w:                          # chained fold (v_proj / q_proj / k_proj carry an earlier fold): tiny, rarely used path
            dWin = dfull
            if hc:
                dWin = dWin * colc.view(1, -1)
            if hr:
                dWin = dWin * (rowc if ctx.row_op == 1 else 1.0 / rowc).view(-1, 1)
            if hs and gmv is not None:
                dWin = dWin + torch.outer(e, shc)
        shp = ctx.shapes
        back = lambda g, s: None if g is None else g.view(s)
        return dWin, back(dcol, shp[0]), back(drow, shp[1]), None, back(dsh, shp[2])


def _fold(W, col=None, row=None, row_op=0, shift=None):
    return _FoldFn.apply(W, col, row, row_op, shift)


def _bias_of(m):
    b = getattr(m, "bias", None)
    return b


def _set_bias_inplace(m, value):
    """in-place variants create the buffer when the module had no bias (transformation.py:79-80,88-89,104-105)."""
    if _bias_of(m) is not None:
        m.bias.copy_(value)
    else:
        if hasattr(m, "bias"):
            del m.bias
        m.register_buffer("bias", value)


def _norm_fold(ln, scales, shifts):
    """O(H) vectors of the norm: torch elementwise (plumbing; the full-size weight work is in _fold)."""
    b = _bias_of(ln)
    new_b = (b - shifts) / scales if b is not None else (-1 * shifts) / scales      # :28-31
    return ln.weight / scales, new_b                                                  # :33


def _fc_in_fold(fc, scales, shifts):
    """input-channel fold of a linear: W * s[col], b + W @ delta (raw W)  (:37-41, :56-60), one kernel pass."""
    b = _bias_of(fc)
    w, mv = _fold(fc.weight, col=scales, shift=shifts)
    return w, (b + mv) if b is not None else mv


def _apply(m, weight, bias, inplace):
    m.use_temporary_parameter = not inplace
    if inplace:
        m.weight.copy_(weight)
        if bias is not None:
            _set_bias_inplace(m, bias)
    else:
        m.temp_weight, m.temp_bias = weight, bias


def _ln_fcs(ln, fcs, scales, shifts, inplace):
    fcs = fcs if isinstance(fcs, list) else [fcs]
    folded = [_fc_in_fold(fc, scales, shifts) for fc in fcs]       # all from the un-modified weights
    w, b = _norm_fold(ln, scales, shifts)
    _apply(ln, w, b, inplace)
    for fc, (fw, fb) in zip(fcs, folded):
        _apply(fc, fw, fb, inplace)


def _fc_fc(fc1, fc2, scales, shifts, inplace):
    """v_proj -> out_proj (:44-60, :93-106): output-channel fold of fc1 on top of whatever it already
    carries, input-channel fold of fc2."""
    if inplace:
        w1, b1 = fc1.weight, fc1.bias
    elif hasattr(fc1, "temp_weight"):
        w1, b1 = fc1.temp_weight, fc1.temp_bias - shifts
    else:
        w1, b1 = fc1.weight, fc1.bias
    if inplace:
        b1 = b1 - shifts
    w2, b2 = _fc_in_fold(fc2, scales, shifts)
    w1f, _ = _fold(w1, row=scales, row_op=2)
    _apply(fc1, w1f, b1 / scales.view(-1), inplace)
    _apply(fc2, w2, b2, inplace)


def _q_k(q_proj, k_proj, scales, inplace):
    src = (lambda m: (m.weight, m.bias)) if inplace else (lambda m: (m.temp_weight, m.temp_bias))
    (wq, bq), (wk, bk) = src(q_proj), src(k_proj)
    _apply(q_proj, _fold(wq, row=scales, row_op=2)[0], bq / scales.view(-1), inplace)      # :66-67
    _apply(k_proj, _fold(wk, row=scales, row_op=1)[0], bk * scales.view(-1), inplace)      # :68-69


# ---------------------------------------------------------------------------- reference entry points
def smooth_ln_fcs_temporary(ln, fcs, scales, shifts):
    _ln_fcs(ln, fcs, scales, shifts, inplace=False)


def smooth_fc_fc_temporary(fc1, fc2, scales, shifts=None):
    _fc_fc(fc1, fc2, scales, shifts, inplace=False)


def smooth_q_k_temporary(q_proj, k_proj, scales):
    _q_k(q_proj, k_proj, scales, inplace=False)


@torch.no_grad()
def smooth_ln_fcs_inplace(ln, fcs, scales, shifts):
    _ln_fcs(ln, fcs, scales, shifts, inplace=True)


@torch.no_grad()
def smooth_fc_fc_inplace(fc1, fc2, scales, shifts=None):
    _fc_fc(fc1, fc2, scales, shifts, inplace=True)


@torch.no_grad()
def smooth_q_k_inplace(q_proj, k_proj, scales):
    _q_k(q_proj, k_proj, scales, inplace=True)
