"""CPU oracle for the OmniQuant block-wise calibration hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``omniquant_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and only as the checker or as
the timed CPU baseline, never as the product path.

It is a functional restatement (plain tensors + dicts, no nn.Module tree) of the
reference algorithm, written against the reference sources cited per function
(paths relative to /root/reference).  It runs on CPU with torch (same ATen CPU
kernels the reference itself runs on, so fp32 results are bit-identical) and
has a pure-numpy twin for the integer-code part of the quantiser.

Parity pinning: the reference has no tests / golden vectors of its own
(SURVEY.md §4), so this oracle is pinned against fixtures generated from the
*live reference modules* imported in the build container:
``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``; checked by
``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

CLIPMIN = 1e-5  # quantize/quantizer.py:10


# --------------------------------------------------------------------------
# quantiser  (quantize/quantizer.py:84-147)
# --------------------------------------------------------------------------
def _ste_round(x: torch.Tensor) -> torch.Tensor:
    """quantize/quantizer.py:15-19 -- round with identity gradient."""
    return (x.round() - x).detach() + x


def _group_view(x: torch.Tensor, group_size: Optional[int]):
    """quantize/quantizer.py:123-129 -- rows of ``group_size`` (zero padded when
    the last dim is not a multiple; the reference only allows that for
    symmetric quantisation, :66-69)."""
    if not group_size:
        return x, 0
    deficiency = (-x.shape[-1]) % group_size
    if deficiency:
        x = torch.cat((x, torch.zeros((x.shape[0], deficiency), dtype=x.dtype, device=x.device)), dim=1)
    return x.reshape(-1, group_size), deficiency


def qparams(x: torch.Tensor, n_bits: int, symmetric: bool = False,
            group_size: Optional[int] = None,
            up: Optional[torch.Tensor] = None, low: Optional[torch.Tensor] = None,
            recip_div: bool = False):
    """scale and rounded zero point, quantize/quantizer.py:122-147.

    ``recip_div`` selects ATen's CUDA semantics for ``tensor / python_scalar``
    (multiply by the fp32 reciprocal) instead of the CPU true division; see
    SURVEY.md §7 hard part 2.  The asymmetric scale clamp is dead code in the
    reference (:144 overwritten by :145) so it is not applied.
    """
    xg, _ = _group_view(x, group_size)
    xmin = xg.amin(-1, keepdim=True)
    xmax = xg.amax(-1, keepdim=True)
    if up is not None:
        xmax = torch.sigmoid(up) * xmax
        xmin = torch.sigmoid(low) * xmin

    def sdiv(t, d):
        if recip_div:
            return t * torch.tensor(1.0 / d, dtype=torch.float32, device=t.device).to(t.dtype)
        return t / d

    if symmetric:
        abs_max = torch.max(xmax.abs(), xmin.abs())
        scale = sdiv(abs_max, 2 ** (n_bits - 1) - 1).clamp(min=CLIPMIN, max=1e4)
        zp = (2 ** (n_bits - 1) - 1) * torch.ones_like(scale)
    else:
        scale = sdiv(xmax - xmin, 2 ** n_bits - 1)
        zp = -(xmin) / scale
    rzp = zp.clamp(min=-1e4, max=1e4).round()
    return scale, rzp


def fake_quant(x: torch.Tensor, n_bits: int, symmetric: bool = False,
               group_size: Optional[int] = None,
               up: Optional[torch.Tensor] = None, low: Optional[torch.Tensor] = None,
               recip_div: bool = False, return_all: bool = False):
    """UniformAffineQuantizer.forward for per_token / per_channel dynamic
    methods, quantize/quantizer.py:108-120 + :84-105."""
    if n_bits >= 16:
        return (x, None, None, None) if return_all else x
    scale, rzp = qparams(x, n_bits, symmetric, group_size, up, low, recip_div)
    shape = x.shape
    xg, deficiency = _group_view(x, group_size)
    if not group_size:
        xg = x
    codes = (_ste_round(xg / scale) + rzp).clamp(0, 2 ** n_bits - 1)
    deq = (codes - rzp) * scale
    if group_size:
        deq = deq.reshape(shape[0], -1)
        if deficiency:
            deq = deq[:, :-deficiency]
    if return_all:
        return deq, scale, rzp, codes
    return deq


def fake_quant_numpy(x: np.ndarray, n_bits: int, symmetric: bool, group_size: Optional[int],
                     sig_up: Optional[np.ndarray], sig_low: Optional[np.ndarray],
                     recip_div: bool = False):
    """numpy fp32 twin of :func:`fake_quant` (forward only) taking the
    already-evaluated sigmoid(up/low) clip multipliers.  Every intermediate is
    rounded to fp32 exactly as the chain of ATen ops in quantizer.py:131-147,
    :93-100 does.  Returns (dequant, scale, zero_point, codes[uint16])."""
    f = np.float32
    x = np.asarray(x, dtype=f)
    rows, cols = x.shape
    xg = x
    deficiency = 0
    if group_size:
        deficiency = (-cols) % group_size
        if deficiency:
            xg = np.concatenate([x, np.zeros((rows, deficiency), f)], axis=1)
        xg = xg.reshape(-1, group_size)
    xmin = xg.min(axis=1, keepdims=True)
    xmax = xg.max(axis=1, keepdims=True)
    if sig_up is not None:
        xmax = (np.asarray(sig_up, f).reshape(-1, 1) * xmax).astype(f)
        xmin = (np.asarray(sig_low, f).reshape(-1, 1) * xmin).astype(f)

    def sdiv(t, d):
        if recip_div:
            return (t * f(1.0 / d)).astype(f)
        return (t / f(d)).astype(f)

    with np.errstate(divide="ignore", invalid="ignore"):
        if symmetric:
            amax = np.maximum(np.abs(xmax), np.abs(xmin))
            scale = np.clip(sdiv(amax, 2 ** (n_bits - 1) - 1), f(CLIPMIN), f(1e4)).astype(f)
            zp = np.full_like(scale, 2 ** (n_bits - 1) - 1)
        else:
            scale = sdiv((xmax - xmin).astype(f), 2 ** n_bits - 1)
            zp = ((-xmin) / scale).astype(f)
        rzp = np.rint(np.clip(zp, f(-1e4), f(1e4))).astype(f)
        t = (np.rint((xg / scale).astype(f)) + rzp).astype(f)
        codes = np.clip(t, f(0), f(2 ** n_bits - 1)).astype(f)
        deq = ((codes - rzp).astype(f) * scale).astype(f)
    if group_size:
        deq = deq.reshape(rows, -1)
        if deficiency:
            deq = deq[:, :cols]
    return deq, scale, rzp, codes


# --------------------------------------------------------------------------
# LET folds  (models/transformation.py)
# --------------------------------------------------------------------------
def truncate_(p: torch.Tensor, thr: float = 1e-2) -> None:
    """models/transformation.py:5-20 applied to ``.data`` as the block
    wrappers do (models/int_llama_layer.py:281-284)."""
    with torch.no_grad():
        small = p.abs() < thr
        p[small] = p[small].sign() * thr


def fold_ln_fcs(ln_w, ln_b, fcs, scales, shifts):
    """models/transformation.py:24-41.  ``fcs`` = list of (W, b|None).
    Returns (ln_w', ln_b', [(W', b')...])."""
    if ln_b is not None:
        nb = (ln_b - shifts) / scales
    else:
        nb = (-1 * shifts) / scales
    nw = ln_w / scales
    out = []
    for W, b in fcs:
        tb = W @ shifts if b is None else b + W @ shifts
        out.append((W * scales.view(1, -1), tb))
    return nw, nb, out


def fold_fc_fc(fc1, fc2, scales, shifts):
    """models/transformation.py:44-60 (v_proj -> out_proj).  fc1 is the already
    folded (W', b') pair of v_proj; fc2 the raw (W, b|None) of out_proj."""
    W1, b1 = fc1
    b1 = (b1 - shifts) / scales.view(-1)
    W1 = W1 / scales.view(-1, 1)
    W2, b2 = fc2
    tb2 = W2 @ shifts if b2 is None else b2 + W2 @ shifts
    return (W1, b1), (W2 * scales.view(1, -1), tb2)


def fold_q_k(q, k, scales):
    """models/transformation.py:63-69."""
    (Wq, bq), (Wk, bk) = q, k
    return (Wq / scales.view(-1, 1), bq / scales.view(-1)), (Wk * scales.view(-1, 1), bk * scales.view(-1))


# --------------------------------------------------------------------------
# norms  (quantize/omni_norm.py)
# --------------------------------------------------------------------------
def rms_norm(x, w, b, eps):
    """quantize/omni_norm.py:52-63."""
    dt = x.dtype
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    h = x * torch.rsqrt(var + eps)
    return (w * h + b).to(dt) if b is not None else (w * h).to(dt)


def layer_norm(x, w, b, eps):
    """quantize/omni_norm.py:26-33."""
    return F.layer_norm(x, (x.shape[-1],), w, b, eps=eps)


# --------------------------------------------------------------------------
# rotary embedding (transformers 4.31 semantics the reference targets,
# models/int_llama_layer.py:70,124-125)
# --------------------------------------------------------------------------
def rope_tables(head_dim: int, T: int, base: float = 10000.0):
    inv = 1.0 / (base ** (torch.arange(0, head_dim, 2).float() / head_dim))
    t = torch.arange(T).float()
    emb = torch.cat((torch.outer(t, inv),) * 2, dim=-1)
    return emb.cos(), emb.sin()  # [T, d]


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def apply_rope(q, k, cos, sin):
    """q,k: [b,h,T,d]; cos/sin: [T,d]."""
    c, s = cos[None, None], sin[None, None]
    return q * c + _rotate_half(q) * s, k * c + _rotate_half(k) * s


def causal_mask(T: int, dtype=torch.float32):
    m = torch.full((T, T), torch.finfo(dtype).min, dtype=dtype)
    return torch.triu(m, diagonal=1)[None, None]


# --------------------------------------------------------------------------
# configuration + parameter containers
# --------------------------------------------------------------------------
class QCfg:
    """The six quantiser-parameter dicts main.py:268-303 builds, flattened."""

    def __init__(self, wbits=4, abits=4, group_size=None, symmetric=False, lwc=True, let=True,
                 recip_div=False, emu16=None, exact_int=None):
        self.wbits, self.abits, self.group_size = wbits, abits, group_size
        self.symmetric, self.lwc, self.let = symmetric, lwc, let
        self.recip_div = recip_div
        # emu16 (e.g. torch.bfloat16): round every tensor the CUDA engine stores in 16 bits (activations,
        # GEMM operands) at the same points, with identity gradient -- SURVEY.md §7 hard part 2b.  None = the
        # reference's fp32 (--deactive_amp) arithmetic.
        self.emu16 = emu16
        # exact_int: with 16-bit storage emulation, a linear whose activation AND (per-channel) weight are quantised to
        # <= 8 bits is contracted from the un-rounded dequantised operands -- the engine multiplies the integer codes
        # exactly (oq_gemm_s8) instead of rounding (code - zp) * scale to bf16 first.  Default: on whenever emu16 is on.
        self.exact_int = (emu16 is not None) if exact_int is None else exact_int


def st(t, cfg):
    """storage rounding of the 16-bit engine (straight-through)."""
    if t is None or cfg.emu16 is None:
        return t
    return t + (t.to(cfg.emu16).to(t.dtype) - t).detach()


def lin_operand_st(cfg, act_quant):
    """storage rounding applied to the two operands of a QuantLinear GEMM (see QCfg.exact_int)."""
    if cfg.emu16 is not None and cfg.exact_int and act_quant and cfg.abits <= 8 and cfg.wbits <= 8 and not cfg.group_size:
        return lambda t: t
    return lambda t: st(t, cfg)


def st_scores(t, cfg):
    """the engine keeps attention scores in fp16 (like the reference's fp16 autocast), everything else in emu16."""
    if cfg.emu16 is None:
        return t
    return t + (t.to(torch.float16).to(t.dtype) - t).detach()


LLAMA_LINEARS = ("q", "k", "v", "o", "gate", "up", "down")
OPT_LINEARS = ("q", "k", "v", "o", "fc1", "fc2")


def lwc_shape(W: torch.Tensor, group_size: Optional[int]) -> int:
    """quantize/quantizer.py:63-71."""
    if group_size:
        return int(W.shape[0] * math.ceil(W.shape[1] / group_size))
    return W.shape[0]


def init_lwc(weights: Dict[str, torch.Tensor], cfg: QCfg, init_value: float = 4.0):
    """quantize/quantizer.py:62-73 -- up/low factors of shape (dim1,1) at 4.0."""
    P = {}
    if cfg.lwc:
        for n, W in weights.items():
            d = lwc_shape(W, cfg.group_size)
            P[f"{n}.up"] = torch.ones((d, 1)) * init_value
            P[f"{n}.low"] = torch.ones((d, 1)) * init_value
    return P


def init_let_scale(act_scale: torch.Tensor, W: torch.Tensor, alpha: float):
    """quantize/omniquant.py:189-191 (note ``max`` not ``abs().max()``)."""
    act = act_scale.clamp(min=1e-5)
    w = W.max(dim=0)[0].clamp(min=1e-5)
    return (act.pow(alpha) / w.pow(1 - alpha)).clamp(min=1e-5)


def wq(W, name, P, cfg: QCfg):
    """weight fake-quant with this weight's LWC factors
    (models/int_llama_layer.py:299-304)."""
    up = P.get(f"{name}.up") if cfg.lwc else None
    low = P.get(f"{name}.low") if cfg.lwc else None
    return fake_quant(W, cfg.wbits, cfg.symmetric, cfg.group_size, up, low, cfg.recip_div)


def aq(x, cfg: QCfg, enabled=True):
    """per-token dynamic activation fake-quant over the last dim
    (main.py:276-281; quantize/int_linear.py:59-60)."""
    if not enabled or cfg.abits >= 16:
        return x
    return fake_quant(x, cfg.abits, False, None, None, None, cfg.recip_div)


# --------------------------------------------------------------------------
# LLaMA block  (models/int_llama_layer.py)
# --------------------------------------------------------------------------
def llama_prepare(Wt: Dict[str, torch.Tensor], P: Dict[str, torch.Tensor], cfg: QCfg, quant=True):
    """smooth_and_quant_temporary, models/int_llama_layer.py:279-307.
    Wt holds raw fp32 weights q,k,v,o,gate,up,down + ln1,ln2.
    Returns dict of temp weights / biases (None where the reference has none)."""
    T = {}
    if cfg.let:
        for k in ("qkv_s", "out_s", "fc1_s", "qkt_s"):
            truncate_(P[k])
        ln1w, ln1b, (q, k, v) = fold_ln_fcs(Wt["ln1"], None, [(Wt["q"], None), (Wt["k"], None), (Wt["v"], None)],
                                            P["qkv_s"], P["qkv_d"])
        ln2w, ln2b, (up, gate) = fold_ln_fcs(Wt["ln2"], None, [(Wt["up"], None), (Wt["gate"], None)],
                                             P["fc1_s"], P["fc1_d"])
        v, o = fold_fc_fc(v, (Wt["o"], None), P["out_s"], P["out_d"])
        q, k = fold_q_k(q, k, P["qkt_s"])
        down = (Wt["down"], None)
        T.update(ln1w=ln1w, ln1b=ln1b, ln2w=ln2w, ln2b=ln2b)
        fcs = dict(q=q, k=k, v=v, o=o, gate=gate, up=up, down=down)
    else:
        T.update(ln1w=Wt["ln1"], ln1b=None, ln2w=Wt["ln2"], ln2b=None)
        fcs = {n: (Wt[n], None) for n in LLAMA_LINEARS}
    for n, (W, b) in fcs.items():
        T[f"{n}.w"] = wq(W, n, P, cfg) if quant else W
        T[f"{n}.b"] = b
    return T


def llama_forward(x, T, n_heads, cos, sin, mask, cfg: QCfg, act_quant=True, eps=1e-6, n_kv_heads=None):
    """QuantLlamaDecoderLayer.forward, models/int_llama_layer.py:213-267 with
    QuantLlamaAttention.forward :103-179 and QuantLlamaMLP.forward :44-45.
    x: [b,T,H]."""
    b, L, H = x.shape
    d = H // n_heads
    n_kv = n_kv_heads or n_heads
    S = lambda t: st(t, cfg)
    So = lin_operand_st(cfg, act_quant)
    lin = lambda t, n, res=None: S((F.linear(So(aq(t, cfg, act_quant)), So(T[f"{n}.w"]), T[f"{n}.b"])) if res is None else
                                   (res + F.linear(So(aq(t, cfg, act_quant)), So(T[f"{n}.w"]), T[f"{n}.b"])))
    a1 = S(rms_norm(x, T["ln1w"], T["ln1b"], eps))
    q = lin(a1, "q").view(b, L, n_heads, d).transpose(1, 2)
    k = lin(a1, "k").view(b, L, n_kv, d).transpose(1, 2)
    v = lin(a1, "v").view(b, L, n_kv, d).transpose(1, 2)
    q, k = apply_rope(q, k, cos[:L], sin[:L])
    if n_kv != n_heads:
        rep = n_heads // n_kv
        k = k[:, :, None].expand(b, n_kv, rep, L, d).reshape(b, n_heads, L, d)
        v = v[:, :, None].expand(b, n_kv, rep, L, d).reshape(b, n_heads, L, d)
    q = S(aq(q, cfg, act_quant))
    k = S(aq(k, cfg, act_quant))
    s = st_scores(torch.matmul(q, k.transpose(2, 3)) / math.sqrt(d), cfg)
    if mask is not None:
        s = s + mask
        s = torch.max(s, torch.tensor(torch.finfo(s.dtype).min, device=s.device))
    p = S(F.softmax(s, dim=-1, dtype=torch.float32).to(q.dtype))
    v = S(aq(v, cfg, act_quant))
    o = S(torch.matmul(p, v).transpose(1, 2).reshape(b, L, H))
    h1 = lin(o, "o", x)
    a2 = S(rms_norm(h1, T["ln2w"], T["ln2b"], eps))
    m = S(F.silu(lin(a2, "gate")) * lin(a2, "up"))
    return lin(m, "down", h1)


def llama_inplace(Wt, P, cfg: QCfg):
    """smooth_and_quant_inplace, models/int_llama_layer.py:315-332: the folded
    and fake-quantised buffers the block keeps after calibration."""
    with torch.no_grad():
        return llama_prepare(Wt, P, cfg, quant=True)


# --------------------------------------------------------------------------
# OPT block  (models/int_opt_layer.py)
# --------------------------------------------------------------------------
def opt_prepare(Wt, P, cfg: QCfg, quant=True):
    """smooth_and_quant_temporary, models/int_opt_layer.py:385-413.  Wt holds
    q,k,v,o,fc1,fc2 weights, '<n>.bias' biases, ln1/ln1.bias, ln2/ln2.bias."""
    T = {}
    if cfg.let:
        for k in ("qkv_s", "out_s", "fc1_s", "qkt_s"):
            truncate_(P[k])
        pair = lambda n: (Wt[n], Wt[f"{n}.bias"])
        ln1w, ln1b, (q, k, v) = fold_ln_fcs(Wt["ln1"], Wt["ln1.bias"], [pair("q"), pair("k"), pair("v")],
                                            P["qkv_s"], P["qkv_d"])
        ln2w, ln2b, (fc1,) = fold_ln_fcs(Wt["ln2"], Wt["ln2.bias"], [pair("fc1")], P["fc1_s"], P["fc1_d"])
        v, o = fold_fc_fc(v, pair("o"), P["out_s"], P["out_d"])
        q, k = fold_q_k(q, k, P["qkt_s"])
        T.update(ln1w=ln1w, ln1b=ln1b, ln2w=ln2w, ln2b=ln2b)
        fcs = dict(q=q, k=k, v=v, o=o, fc1=fc1, fc2=pair("fc2"))
    else:
        T.update(ln1w=Wt["ln1"], ln1b=Wt["ln1.bias"], ln2w=Wt["ln2"], ln2b=Wt["ln2.bias"])
        fcs = {n: (Wt[n], Wt[f"{n}.bias"]) for n in OPT_LINEARS}
    for n, (W, b) in fcs.items():
        T[f"{n}.w"] = wq(W, n, P, cfg) if quant else W
        T[f"{n}.b"] = b
    return T


def opt_forward(x, T, n_heads, mask, cfg: QCfg, act_quant=True, eps=1e-5):
    """QuantOPTDecoderLayer.forward (pre-LN), models/int_opt_layer.py:268-346
    with QuantOPTAttention.forward :77-213: q is scaled then quantised per
    token over the full hidden dim (:96-97), k and v likewise before the head
    split (:124-130)."""
    b, L, H = x.shape
    d = H // n_heads
    S = lambda t: st(t, cfg)
    So = lin_operand_st(cfg, act_quant)
    lin = lambda t, n, res=None, sc=1.0: S((F.linear(So(aq(t, cfg, act_quant)), So(T[f"{n}.w"]), T[f"{n}.b"]) * sc) if res is None
                                           else (res + F.linear(So(aq(t, cfg, act_quant)), So(T[f"{n}.w"]), T[f"{n}.b"])))
    sh = lambda t: t.view(b, L, n_heads, d).transpose(1, 2).contiguous().view(b * n_heads, L, d)
    a1 = S(layer_norm(x, T["ln1w"], T["ln1b"], eps))
    q = S(aq(lin(a1, "q", sc=d ** -0.5), cfg, act_quant))
    k = S(aq(lin(a1, "k"), cfg, act_quant))
    v = S(aq(lin(a1, "v"), cfg, act_quant))
    s = st_scores(torch.bmm(sh(q), sh(k).transpose(1, 2)), cfg)
    if mask is not None:
        s = s.view(b, n_heads, L, L) + mask
        s = torch.max(s, torch.tensor(torch.finfo(s.dtype).min, device=s.device)).view(b * n_heads, L, L)
    p = S(F.softmax(s, dim=-1))
    o = S(torch.bmm(p, sh(v)).view(b, n_heads, L, d).transpose(1, 2).reshape(b, L, H))
    h1 = lin(o, "o", x)
    r = h1.reshape(-1, H)
    a2 = S(layer_norm(r, T["ln2w"], T["ln2b"], eps))
    f = lin(S(F.relu(lin(a2, "fc1"))), "fc2", r)
    return f.view(b, L, H)


# --------------------------------------------------------------------------
# Falcon block, new decoder architecture (models/int_falcon_layer.py) -- LWC only (:329-331,:344-346 raise on LET)
# --------------------------------------------------------------------------
FALCON_LINEARS = ("qkv", "dense", "fc1", "fc2")


def falcon_prepare(Wt, P, cfg: QCfg, quant=True):
    """smooth_and_quant_temporary, models/int_falcon_layer.py:344-360.  Wt: qkv, dense, fc1, fc2 weights
    (+ optional '<n>.bias'), ln1/ln1.bias (ln_attn), ln2/ln2.bias (ln_mlp)."""
    T = dict(ln1w=Wt["ln1"], ln1b=Wt["ln1.bias"], ln2w=Wt["ln2"], ln2b=Wt["ln2.bias"])
    for n in FALCON_LINEARS:
        T[f"{n}.w"] = wq(Wt[n], n, P, cfg) if quant else Wt[n]
        T[f"{n}.b"] = Wt.get(f"{n}.bias")
    return T


def falcon_forward(x, T, n_heads, n_kv, cos, sin, cfg: QCfg, act_quant=True, eps=1e-5):
    """QuantFalconDecoderLayer.forward with new_decoder_architecture (models/int_falcon_layer.py:261-316) and
    QuantFalconAttention.forward (:122-180): fused QKV laid out per K/V group as [rep query heads | k | v] (:76-86),
    rotary on q and k, causal softmax(q k^T / sqrt(d)) v without any Q/K/V quantisation, attention and MLP in
    parallel on two LayerNorms of the block input, exact (erf) GELU."""
    b, L, H = x.shape
    d = H // n_heads
    rep = n_heads // n_kv
    S = lambda t: st(t, cfg)
    So = lin_operand_st(cfg, act_quant)
    lin = lambda t, n, res=None: S(F.linear(So(aq(t, cfg, act_quant)), So(T[f"{n}.w"]), T[f"{n}.b"]) if res is None else
                                   res + F.linear(So(aq(t, cfg, act_quant)), So(T[f"{n}.w"]), T[f"{n}.b"]))
    a1 = S(layer_norm(x, T["ln1w"], T["ln1b"], eps))
    a2 = S(layer_norm(x, T["ln2w"], T["ln2b"], eps))
    qkv = lin(a1, "qkv").view(b, L, n_kv, rep + 2, d)
    q = qkv[:, :, :, :-2].reshape(b, L, n_heads, d).transpose(1, 2)
    k = qkv[:, :, :, -2].transpose(1, 2)                     # [b, n_kv, L, d]
    v = qkv[:, :, :, -1].transpose(1, 2)
    q, k = apply_rope(q, k, cos[:L], sin[:L])
    q, k = S(q), S(k)
    k = k[:, :, None].expand(b, n_kv, rep, L, d).reshape(b, n_heads, L, d)
    v = v[:, :, None].expand(b, n_kv, rep, L, d).reshape(b, n_heads, L, d)
    s = st_scores(torch.matmul(q, k.transpose(2, 3)) / math.sqrt(d), cfg)
    causal = torch.triu(torch.ones(L, L, dtype=torch.bool, device=x.device), diagonal=1)
    s = s + torch.zeros_like(s).masked_fill(causal, -1e9)
    p = S(F.softmax(s, dim=-1))
    o = S(torch.matmul(p, v).transpose(1, 2).reshape(b, L, H))
    h1 = lin(o, "dense", x)
    m = S(F.gelu(lin(a2, "fc1")))
    return lin(m, "fc2", h1)


# --------------------------------------------------------------------------
# calibration loop  (quantize/omniquant.py:157-250)
# --------------------------------------------------------------------------
def calibrate_block(family: str, Wt, P, cfg: QCfg, inps, n_heads, *, epochs, batch_size=1,
                    let_lr=5e-3, lwc_lr=1e-2, wd=0.0, aug_loss=False, eps=None, steps_limit=None,
                    n_kv_heads=None, record_grads=False, quant_inps=None):
    """One decoder block of omniquant(): FP pass (:165-172), AdamW loop
    (:207-234), in-place fold + quantise (:239), requant pass (:242-245).
    fp32 throughout (the reference's ``--deactive_amp`` arithmetic).  Returns
    dict(losses, P (trained), T_final, quant_out, fp_out, grads)."""
    nsamples, L, H = inps.shape
    d = H // n_heads
    dev = inps.device
    mask = causal_mask(L).to(dev)
    if family == "llama":
        cos, sin = rope_tables(d, L)
        cos, sin = cos.to(dev), sin.to(dev)
        eps = 1e-6 if eps is None else eps
        prep = llama_prepare
        fwd = lambda x, T, aqf: llama_forward(x, T, n_heads, cos, sin, mask, cfg, aqf, eps, n_kv_heads)
    elif family == "falcon":
        cos, sin = rope_tables(d, L)
        cos, sin = cos.to(dev), sin.to(dev)
        eps = 1e-5 if eps is None else eps
        prep = falcon_prepare
        fwd = lambda x, T, aqf: falcon_forward(x, T, n_heads, n_kv_heads or n_heads, cos, sin, cfg, aqf, eps)
    else:
        eps = 1e-5 if eps is None else eps
        prep = opt_prepare
        fwd = lambda x, T, aqf: opt_forward(x, T, n_heads, mask, cfg, aqf, eps)
    no_let = QCfg(cfg.wbits, cfg.abits, cfg.group_size, cfg.symmetric, cfg.lwc, False, cfg.recip_div, cfg.emu16)
    with torch.no_grad():
        T_fp = prep(Wt, P, no_let, quant=False)
        fp_out = torch.stack([fwd(inps[j:j + 1], T_fp, False)[0] for j in range(nsamples)])
    quant_inps = inps.clone() if quant_inps is None else quant_inps.clone()
    fp_out_2 = None
    if aug_loss:  # quantize/omniquant.py:171-172: FP block applied to the quantised stream
        with torch.no_grad():
            fp_out_2 = torch.stack([fwd(quant_inps[j:j + 1], T_fp, False)[0] for j in range(nsamples)])
    let_keys = [k for k in P if k.endswith("_s") or k.endswith("_d")]
    lwc_keys = [k for k in P if k.endswith(".up") or k.endswith(".low")]
    for k in let_keys + lwc_keys:
        P[k].requires_grad_(True)
    losses, grads0 = [], None
    if epochs > 0:
        opt = torch.optim.AdamW([{"params": [P[k] for k in let_keys], "lr": let_lr},
                                 {"params": [P[k] for k in lwc_keys], "lr": lwc_lr}], weight_decay=wd)
        step = 0
        for _ in range(epochs):
            for j in range(nsamples // batch_size):
                i0 = j * batch_size
                T = prep(Wt, P, cfg, quant=True)
                out = fwd(quant_inps[i0:i0 + batch_size], T, True)
                loss = F.mse_loss(fp_out[i0:i0 + batch_size], out)
                if aug_loss:
                    loss = loss + F.mse_loss(fp_out_2[i0:i0 + batch_size], out)
                losses.append(float(loss.detach()))
                opt.zero_grad()
                loss.backward()
                if record_grads and grads0 is None:
                    grads0 = {k: P[k].grad.detach().clone() for k in let_keys + lwc_keys if P[k].grad is not None}
                opt.step()
                step += 1
                if steps_limit and step >= steps_limit:
                    break
            if steps_limit and step >= steps_limit:
                break
    with torch.no_grad():
        T_final = prep(Wt, P, cfg, quant=True)
        qout = torch.stack([fwd(quant_inps[j:j + 1], T_final, True)[0] for j in range(nsamples)])
    for k in let_keys + lwc_keys:
        P[k].requires_grad_(False)
    return dict(losses=losses, P=P, T_final=T_final, quant_out=qout, fp_out=fp_out, grads=grads0)


# ---------------------------------------------------------------- activation statistics (LET initialisation inputs)
def act_stats_sample(x):
    """One sample's per-channel statistics of a linear layer's input x [..., C]: (max|x|, (max+min)/2), both fp32
    (generate_act_scale_shift.py:30-33 stat_tensor of get_act_scales; :64-67,70,72 of get_act_shifts)."""
    t = x.reshape(-1, x.shape[-1])
    absmax = t.abs().max(dim=0)[0].float()
    mid = (t.max(dim=0)[0].float() + t.min(dim=0)[0].float()) / 2
    return absmax, mid


def act_stats_fold(absmaxs, mids):
    """Fold per-sample vectors in sample order: running max (generate_act_scale_shift.py:34-37) and the EMA
    0.99*prev + 0.01*mid started from the first sample (:69-72)."""
    scale, shift = absmaxs[0].clone(), mids[0].clone()
    for i in range(1, len(absmaxs)):
        scale = torch.max(scale, absmaxs[i])
        shift = 0.99 * shift + 0.01 * mids[i]
    return scale, shift
