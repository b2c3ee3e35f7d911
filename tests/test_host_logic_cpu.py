"""Host-side logic that needs no GPU: the attention head-pass partition of the block engine, the block streamer's
synchronous fall-back, the perplexity head lookup and the ABI surface of the fused entry points."""
import os
from types import SimpleNamespace

import pytest
import torch
from torch import nn


def _fake_engine(h, hkv, T, d, family="llama", causal=True):
    from omniquant_b200.engine import BlockEngine
    e = SimpleNamespace(h=h, hkv=hkv, rep=h // hkv, T=T, d=d, H=h * d, Hkv=hkv * d, family=family, causal=causal, _passes=None)
    e._att_layout = lambda: BlockEngine._att_layout(e)
    return e, BlockEngine._att_passes


@pytest.mark.parametrize("h,hkv,budget_heads", [(32, 32, 8), (32, 32, 5), (64, 8, 16), (8, 2, 4), (12, 12, 0)])
def test_attention_passes_partition_heads(monkeypatch, h, hkv, budget_heads):
    T, d = 2048, 128
    per_head = T * T * 4 * 0.5
    monkeypatch.setenv("OQ_ATT_PASS_MB", str(budget_heads * per_head / 1e6))
    e, passes_of = _fake_engine(h, hkv, T, d)
    passes = passes_of(e)
    heads = []
    for c0, c1, groups in passes:
        assert c0 == groups[0][0] and c1 == groups[-1][0] + groups[-1][1]
        for h0, nh, cq, ck, cv, sb in groups:
            heads += list(range(h0, h0 + nh))
            if hkv == h:                                     # MHA: q / k / v column offsets follow the head index
                assert (cq, ck, cv) == (h0 * d, h * d + h0 * d, 2 * h * d + h0 * d)
    assert heads == list(range(h))                           # every head exactly once, in order
    if budget_heads == 0:
        assert len(passes) == 1
    else:
        rep = h // hkv
        assert all(c1 - c0 <= max(budget_heads, rep) for c0, c1, _ in passes)
        if hkv != h:
            assert all(c0 % rep == 0 and c1 % rep == 0 for c0, c1, _ in passes)   # whole K/V groups per pass


def test_block_streamer_cpu_fallback():
    from omniquant_b200.streaming import BlockStreamer
    layers = nn.ModuleList([nn.Linear(4, 4) for _ in range(3)])
    s = BlockStreamer(layers, "cpu")
    assert not s.enabled
    s.prefetch(1)                                            # no-op without a GPU
    assert s._pending == {}
    assert s.fetch(1) is layers[1]
    assert s.bytes_of(0) == (16 + 4) * 4


def test_eval_head_lookup():
    from omniquant_b200.evaluate import _head
    from omniquant_b200.main import synthetic_model
    m = synthetic_model("llama-tiny")
    norm, proj, head = _head(m, "llama")
    assert norm is m.model.norm and proj is None and head is m.lm_head
    m = synthetic_model("opt-tiny")
    norm, proj, head = _head(m, "opt")
    assert head is m.lm_head and norm is m.model.decoder.final_layer_norm


def test_fused_entry_points_are_declared_and_exported():
    from omniquant_b200 import _lib
    names = set(_lib.declared_functions())
    for fn in ("oq_norm_act_quant_fwd", "oq_swiglu_act_quant_fwd", "oq_swiglu_act_quant_bwd"):
        assert fn in names
    sig = _lib.lib().fns["oq_act_quant_fwd"]
    assert sig is not None
    hdr = open(os.path.join(os.path.dirname(_lib.__file__), "..", "include", "omniquant_b200.h")).read()
    assert "int rope_inner" in hdr                            # one launch covers q | k (rotated) | v (not rotated)


def test_integer_forward_gemm_identity():
    """DESIGN.md section 9, item 1 (numerics of the planned exact-integer forward GEMM): with both operands fake-quantised,
    Y = s_a[m] * s_w[n] * sum_k (c_a - z_a)(c_w - z_w) with the integer sum accumulated exactly; it reproduces the fp64 product
    of the oracle's dequantised operands to fp32 round-off and is closer to it than the product of bf16-rounded operands."""
    from oracle import oq_oracle as orc
    g = torch.Generator().manual_seed(3)
    M, K, N = 24, 512, 40
    x = torch.randn(M, K, generator=g) * 2
    W = torch.randn(N, K, generator=g) * 0.02
    for bits in (4, 6):
        xq, sa, za, ca = orc.fake_quant(x, bits, return_all=True)                     # per token
        up = torch.full((N, 1), 4.0)
        wq, sw, zw, cw = orc.fake_quant(W, bits, False, None, up, up.clone(), return_all=True)   # per channel + LWC
        ia = (ca.reshape(M, K) - za.reshape(M, 1)).to(torch.int64)
        iw = (cw.reshape(N, K) - zw.reshape(N, 1)).to(torch.int64)
        assert int(ia.abs().max()) < 128 and int(iw.abs().max()) < 128               # s8 operands
        acc = ia @ iw.t()                                                             # exact
        assert int(acc.abs().max()) < 2 ** 24                                         # also exact in an fp32 accumulator
        y_int = acc.double() * sa.reshape(M, 1).double() * sw.reshape(1, N).double()
        y_ref = xq.double() @ wq.double().t()
        y_bf = (xq.bfloat16().double()) @ (wq.bfloat16().double()).t()
        e_int = float((y_int - y_ref).abs().max() / y_ref.abs().max())
        e_bf = float((y_bf - y_ref).abs().max() / y_ref.abs().max())
        assert e_int < 1e-6 and e_int < e_bf / 50, (bits, e_int, e_bf)
