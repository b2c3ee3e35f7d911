"""Sub-block-level attention forward of the three families on the library's kernels: what
``QuantLlamaAttention.forward`` (reference models/int_llama_layer.py:103-187), ``QuantOPTAttention.forward``
(models/int_opt_layer.py:81-197) and ``QuantFalconAttention.forward`` (models/int_falcon_layer.py:126-176) compute for a
causal sequence without a KV cache.  The calibration path never comes through here (the block engine schedules the whole
decoder block, engine.py); these forwards make the attention modules usable on their own, with autograd.

The projections are the QuantLinear modules, the per-token / per-head quantisers the QuantMatMul modules' own; the
contraction QK^T -> causal fp32 softmax -> PV is ONE fused tcgen05 kernel (``oq_attn_fwd``, recompute backward
``oq_attn_bwd``) whenever the shape allows it and P is not quantised, else the two QuantMatMul GEMMs around a torch softmax.
"""
import math
from typing import Optional

import torch

from .. import ops
from .._lib import OqError

BF16 = torch.bfloat16


class _CausalAttnFn(torch.autograd.Function):
    """O[b, T, h*d] from the fused operand matrix qkv [b*T, (h + 2 hkv) * d] = [q heads | k heads | v heads]."""

    @staticmethod
    def forward(ctx, qkv, batch, T, heads, kv_heads, d, alpha):
        M, H, Hkv = batch * T, heads * d, kv_heads * d
        layout = dict(cq=0, gq=(heads // kv_heads) * d, ck=H, cv=H + Hkv, gk=d)
        O = torch.empty(M, H, device=qkv.device, dtype=BF16)
        stats = (torch.empty(batch, heads, T, device=qkv.device), torch.empty(batch, heads, T, device=qkv.device))
        kw = dict(batch=batch, T=T, heads=heads, kv_heads=kv_heads, head_dim=d, layout=layout, alpha=alpha)
        ops.attn_fwd(qkv, O, *stats, **kw)
        ctx.save_for_backward(qkv, O, *stats)
        ctx.kw = kw
        return O

    @staticmethod
    def backward(ctx, dO):
        qkv, O, rmax, rinv = ctx.saved_tensors
        dqkv = torch.empty_like(qkv)
        ops.attn_bwd(qkv, O, dO.to(BF16).contiguous(), rmax, rinv, torch.empty_like(rmax), dqkv, **ctx.kw)
        return dqkv, None, None, None, None, None, None


def rope_tables(head_dim: int, positions: torch.Tensor, theta: float = 10000.0):
    """cos / sin [T, d] of LlamaRotaryEmbedding (transformers 4.31, rotate_half convention) at the given positions."""
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32, device=positions.device) / head_dim))
    emb = torch.outer(positions.float(), inv)
    emb = torch.cat((emb, emb), -1)
    return emb.cos(), emb.sin()


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor):
    """x [b, heads, T, d]: x * cos + rotate_half(x) * sin."""
    d2 = x.shape[-1] // 2
    rot = torch.cat((-x[..., d2:], x[..., :d2]), -1)
    return (x.float() * cos + rot.float() * sin).to(x.dtype)


def check_causal(attention_mask: Optional[torch.Tensor], T: int, additive: bool = True):
    """The reference's callers always hand the causal mask of the calibration sequence; anything else is not implemented."""
    if attention_mask is None or T == 1:
        return
    m = attention_mask.reshape(-1, T, T)[0]
    upper_masked = bool(m[0, T - 1] < -1e4) if additive else bool(m[0, T - 1])
    lower_open = bool(m[T - 1, 0] == 0) if additive else not bool(m[T - 1, 0])
    if not (upper_masked and lower_open):
        raise NotImplementedError("attention_mask must be the causal mask (no padding / prefix masks)")


def causal_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, alpha: float, pv_matmul=None, qkt_matmul=None):
    """q [b, h, T, d], k / v [b, hkv, T, d] (already rotated / fake-quantised) -> [b, T, h * d]."""
    b, h, T, d = q.shape
    hkv = k.shape[1]
    if not q.is_cuda:
        raise OqError("attention forward runs on the CUDA kernels only")
    p_quant = pv_matmul is not None and pv_matmul.use_act_quant and pv_matmul.x1_quantizer.n_bits < 16
    if ops.attn_supported(T, d) and not p_quant:
        pack = lambda t, n: t.transpose(1, 2).reshape(b * T, n * d)
        qkv = torch.cat((pack(q, h), pack(k, hkv), pack(v, hkv)), -1).to(BF16).contiguous()
        out = _CausalAttnFn.apply(qkv, b, T, h, hkv, d, float(alpha))
        return out.view(b, T, h * d).to(q.dtype)
    # two GEMMs around a torch softmax (any T / head_dim, quantised P)
    from ..quantize.int_matmul import quant_bmm
    rep = h // hkv
    if rep > 1:
        k, v = k.repeat_interleave(rep, 1), v.repeat_interleave(rep, 1)
    s = quant_bmm(q, k.transpose(2, 3)) * alpha
    mask = torch.triu(torch.full((T, T), torch.finfo(torch.float32).min, device=q.device), diagonal=1)
    p = torch.softmax(s.float() + mask, dim=-1).to(q.dtype)
    if pv_matmul is not None:
        p = pv_matmul.quant_x1(p)
    out = quant_bmm(p, v)
    return out.transpose(1, 2).reshape(b, T, h * d)


def llama_attention_forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value=None,
                            output_attentions=False, use_cache=False, rope_theta=10000.0):
    if past_key_value is not None or use_cache or output_attentions:
        raise NotImplementedError("KV cache / attention weights output: calibration and evaluation run whole sequences")
    b, T, _ = hidden_states.shape
    h, hkv, d = self.num_heads, self.num_key_value_heads, self.head_dim
    check_causal(attention_mask, T)
    split = lambda t, n: t.view(b, T, n, d).transpose(1, 2)
    q, k, v = split(self.q_proj(hidden_states), h), split(self.k_proj(hidden_states), hkv), split(self.v_proj(hidden_states), hkv)
    pos = position_ids.reshape(-1, T)[0] if position_ids is not None else torch.arange(T, device=q.device)
    cos, sin = rope_tables(d, pos, rope_theta)
    q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
    q, k = self.qkt_matmul.quant_x1(q), self.qkt_matmul.quant_x2(k)      # per (token, head) over head_dim (:141-142)
    v = self.pv_matmul.quant_x2(v)                                        # (:162)
    out = causal_attention(q, k, v, 1.0 / math.sqrt(d), self.pv_matmul, self.qkt_matmul)
    return self.o_proj(out), None, None


def opt_attention_forward(self, hidden_states, key_value_states=None, past_key_value=None, attention_mask=None,
                          layer_head_mask=None, output_attentions=False):
    if key_value_states is not None or past_key_value is not None or layer_head_mask is not None or output_attentions:
        raise NotImplementedError("cross attention / KV cache / head masks / attention weights output")
    b, T, _ = hidden_states.shape
    h, d = self.num_heads, self.head_dim
    check_causal(attention_mask, T)
    q = self.qkt_matmul.quant_x1(self.q_proj(hidden_states) * self.scaling)     # per token over the hidden dim (:96-97)
    k = self.qkt_matmul.quant_x2(self.k_proj(hidden_states))                    # (:124-127)
    v = self.pv_matmul.quant_x2(self.v_proj(hidden_states))                     # (:128-130)
    split = lambda t: t.view(b, T, h, d).transpose(1, 2)
    out = causal_attention(split(q), split(k), split(v), 1.0, self.pv_matmul, self.qkt_matmul)
    return self.out_proj(out), None, None


def falcon_attention_forward(self, hidden_states, alibi, attention_mask, layer_past=None, head_mask=None, use_cache=False,
                             output_attentions=False, rope_theta=10000.0):
    if alibi is not None or layer_past is not None or head_mask is not None or use_cache or output_attentions:
        raise NotImplementedError("alibi / KV cache / head masks / attention weights output")
    b, T, _ = hidden_states.shape
    h, hkv, d = self.num_heads, self.num_kv_heads, self.head_dim
    check_causal(attention_mask, T, additive=attention_mask is not None and attention_mask.dtype != torch.bool)
    rep = h // hkv
    fused = self.query_key_value(hidden_states).view(b, T, hkv, rep + 2, d)     # [rep q heads | k | v] per K/V group (:84-93)
    q = fused[:, :, :, :rep].reshape(b, T, h, d).transpose(1, 2)
    k, v = fused[:, :, :, rep].transpose(1, 2), fused[:, :, :, rep + 1].transpose(1, 2)
    cos, sin = rope_tables(d, torch.arange(T, device=q.device), rope_theta)
    q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
    out = causal_attention(q, k, v, self.inv_norm_factor)
    return self.dense(out), None
