"""The reference's module API as a drop-in: the training loop of quantize/omniquant.py:203-250 written
against the block wrapper (smooth_and_quant_temporary -> forward -> MSELoss -> backward -> AdamW.step ->
smooth_and_quant_inplace), plus the stand-alone operators, all on the CUDA kernels."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from golden_util import BlockFixture, LLAMA_MAP

pytestmark = pytest.mark.gpu


def make_args(fx):
    a = SimpleNamespace(wbits=fx.wbits, abits=fx.abits, group_size=fx.group_size, lwc=fx.lwc, let=fx.let, symmetric=False,
                        let_lr=5e-3, lwc_lr=1e-2, wd=0.0, alpha=0.5, recip_div=False)
    a.weight_quant_params = dict(n_bits=fx.wbits, per_channel_axes=[0], symmetric=False, dynamic_method="per_channel",
                                 group_size=fx.group_size, lwc=fx.lwc)
    tok = dict(n_bits=fx.abits, per_channel_axes=[], symmetric=False, dynamic_method="per_token")
    a.act_quant_params, a.q_quant_params, a.k_quant_params, a.v_quant_params = dict(tok), dict(tok), dict(tok), dict(tok)
    a.p_quant_params = dict(n_bits=16, metric="fix0to1")
    return a


def hf_llama_layer(fx):
    from transformers import LlamaConfig
    from transformers.models.llama.modeling_llama import LlamaDecoderLayer
    cfg = LlamaConfig(hidden_size=fx.H, intermediate_size=fx.F, num_attention_heads=fx.heads, num_key_value_heads=fx.heads,
                      num_hidden_layers=1, vocab_size=64, max_position_embeddings=2048, rms_norm_eps=1e-6)
    layer = LlamaDecoderLayer(cfg, 0)
    W = fx.weights()
    with torch.no_grad():
        for s, long in LLAMA_MAP.items():
            m = layer
            for p in long.split("."):
                m = getattr(m, p)
            m.weight.copy_(W[s])
        layer.input_layernorm.weight.copy_(W["ln1"])
        layer.post_attention_layernorm.weight.copy_(W["ln2"])
    return cfg, layer.half().cuda()


@pytest.mark.parametrize("fname", ["llama_w4a16g64_lwc_let.npz", "llama_w4a4_lwc_let.npz"])
def test_reference_style_loop_on_block_wrapper(fname):
    from omniquant_b200.models.int_llama_layer import QuantLlamaDecoderLayer
    from omniquant_b200.quantize.int_linear import QuantLinear
    fx = BlockFixture(fname)
    args = make_args(fx)
    cfg, layer = hf_llama_layer(fx)
    qlayer = QuantLlamaDecoderLayer(cfg, layer, args).cuda()
    x = fx.inps.cuda().to(torch.bfloat16)
    T = fx.T
    mask = torch.triu(torch.full((T, T), torch.finfo(torch.float32).min, device="cuda"), diagonal=1)[None, None]
    pos = torch.arange(T, device="cuda")[None]
    # FP targets
    qlayer.set_quant_state(weight_quant=False, act_quant=False)
    with torch.no_grad():
        fp = torch.stack([qlayer(x[j:j + 1], attention_mask=mask, position_ids=pos)[0][0] for j in range(fx.nsamples)])
    assert float((fp.float().cpu() - fx.fp_out).norm() / fx.fp_out.norm()) < 1e-2
    qlayer.set_quant_state(weight_quant=False, act_quant=True)
    qlayer.let = True
    init = fx.init_params()
    for short, long in (("qkv_s", "qkv_smooth_scale"), ("qkv_d", "qkv_smooth_shift"), ("out_s", "out_smooth_scale"),
                        ("out_d", "out_smooth_shift"), ("fc1_s", "fc1_smooth_scale"), ("fc1_d", "fc1_smooth_shift"),
                        ("qkt_s", "qkt_smooth_scale")):
        qlayer.register_parameter(long, torch.nn.Parameter(init[short].cuda()))
    qlayer.float()
    opt = torch.optim.AdamW([{"params": qlayer.let_parameters(True), "lr": 5e-3},
                             {"params": qlayer.lwc_parameters(), "lr": 1e-2}], weight_decay=0)
    assert len(list(qlayer.omni_parameters())) == 7 + 14
    keys = list(qlayer.omni_state_dict().keys())
    assert "qkv_smooth_scale" in keys and "self_attn.q_proj.weight_quantizer.upbound_factor" in keys
    loss_fn = torch.nn.MSELoss()
    losses = []
    for _ in range(fx.epochs):
        for j in range(fx.nsamples):
            qlayer.smooth_and_quant_temporary(seq_len=T, batch=1)
            out = qlayer(x[j:j + 1], attention_mask=mask, position_ids=pos)[0]
            loss = loss_fn(fp[j:j + 1].float(), out.float())
            losses.append(float(loss))
            opt.zero_grad()
            loss.backward()
            opt.step()
    ref = fx.losses
    rel = np.abs(np.array(losses) - ref) / ref
    print(fname, "module-loop losses", losses[:4], "ref", ref[:4], "max rel", rel.max())
    assert rel.max() < 8e-2 and abs(np.mean(losses) - ref.mean()) / ref.mean() < 3e-2
    qlayer.clear_temp_variable()
    qlayer.smooth_and_quant_inplace()
    with torch.no_grad():
        qo = torch.stack([qlayer(x[j:j + 1], attention_mask=mask, position_ids=pos)[0][0] for j in range(fx.nsamples)])
    assert float((qo.float().cpu() - fx.quant_out).norm() / fx.quant_out.norm()) < 8e-2
    qlayer.register_scales_and_zeros()
    lin = {n: m for n, m in qlayer.named_modules() if isinstance(m, QuantLinear)}
    assert len(lin) == 7
    Q = fx.quantized()
    w = lin["self_attn.q_proj"].weight.float().cpu()
    assert float((w - Q["q.w"]).abs().mean() / Q["q.w"].abs().mean()) < 5e-2
    assert lin["self_attn.q_proj"].weight_quantizer.scales.shape == Q["q.scales"].shape
    assert qlayer.input_layernorm.bias is not None and lin["self_attn.q_proj"].bias is not None   # folded LET buffers
    # few levels per row/group after fake-quant
    g = fx.group_size or w.shape[1]
    assert all(len(torch.unique(row)) <= 2 ** fx.wbits for row in w.reshape(-1, g)[:16])


def test_standalone_operators():
    from omniquant_b200.quantize.int_linear import QuantLinear
    from omniquant_b200.quantize.int_matmul import QuantMatMul
    from omniquant_b200.quantize.omni_norm import OmniLayerNorm, OmniLlamaRMSNorm
    from omniquant_b200.quantize.quantizer import UniformAffineQuantizer
    from oracle import oq_oracle as orc
    torch.manual_seed(0)
    lin = torch.nn.Linear(256, 384).cuda()
    wq = dict(n_bits=4, per_channel_axes=[0], symmetric=False, dynamic_method="per_channel", group_size=64, lwc=True)
    aq = dict(n_bits=4, per_channel_axes=[], symmetric=False, dynamic_method="per_token")
    ql = QuantLinear(lin, wq, aq).cuda()
    ql.set_quant_state(weight_quant=True, act_quant=True)
    x = torch.randn(2, 40, 256, device="cuda").to(torch.bfloat16)
    y = ql(x)
    with torch.no_grad():
        w_ref = orc.fake_quant(lin.weight.detach().float(), 4, False, 64, ql.weight_quantizer.upbound_factor.detach(),
                               ql.weight_quantizer.lowbound_factor.detach())
        x_ref = orc.fake_quant(x.float(), 4)
        y_ref = torch.nn.functional.linear(x_ref, w_ref, lin.bias.float())
    torch.testing.assert_close(y.float(), y_ref, rtol=3e-2, atol=3e-2)
    y.float().square().mean().backward()
    assert ql.weight_quantizer.upbound_factor.grad is not None and torch.isfinite(ql.weight_quantizer.upbound_factor.grad).all()
    # quantiser: identity cases
    p = UniformAffineQuantizer(n_bits=16, metric="fix0to1")
    t = torch.rand(4, 4, device="cuda")
    assert p(t) is t
    with pytest.raises(Exception):
        UniformAffineQuantizer(**aq)(torch.randn(4, 8))          # CPU tensor: no fallback
    # matmul + norms
    mm = QuantMatMul(aq, aq, matmul_func=torch.bmm).cuda()
    mm.set_quant_state(False, True)
    a, b = torch.randn(3, 64, 128, device="cuda").to(torch.bfloat16), torch.randn(3, 128, 72, device="cuda").to(torch.bfloat16)
    out = mm(mm.quant_x1(a), b)
    ref = torch.bmm(orc.fake_quant(a.float(), 4), b.float())
    torch.testing.assert_close(out.float(), ref, rtol=3e-2, atol=1e-1)
    ln = torch.nn.LayerNorm(256).cuda()
    xs = torch.randn(8, 256, device="cuda")
    torch.testing.assert_close(OmniLayerNorm(ln).cuda()(xs), ln(xs), rtol=2e-2, atol=2e-2)
    rn = SimpleNamespace(weight=torch.nn.Parameter(torch.rand(256, device="cuda") + 0.5))
    torch.testing.assert_close(OmniLlamaRMSNorm(rn, eps=1e-6).cuda()(xs), orc.rms_norm(xs, rn.weight.detach(), None, 1e-6),
                               rtol=2e-2, atol=2e-2)


def test_falcon_block_wrapper_matches_reference_fixture():
    """QuantFalconDecoderLayer (new decoder architecture, LWC only) built from a HF FalconDecoderLayer holding the
    fixture's weights: FP targets, the module-level training loop and the in-place quantised weights."""
    from transformers import FalconConfig
    from transformers.models.falcon.modeling_falcon import FalconDecoderLayer
    from golden_util import FALCON_MAP
    from omniquant_b200.models.int_falcon_layer import QuantFalconDecoderLayer
    from omniquant_b200.quantize.int_linear import QuantLinear
    fx = BlockFixture("falcon_w3a16g64_lwc.npz")
    args = make_args(fx)
    cfg = FalconConfig(hidden_size=fx.H, num_attention_heads=fx.heads, num_kv_heads=fx.kv_heads, num_hidden_layers=1,
                       vocab_size=64, new_decoder_architecture=True, bias=False, parallel_attn=True, ffn_hidden_size=fx.F)
    cfg._attn_implementation = "eager"
    layer = FalconDecoderLayer(cfg, 0)
    W = fx.weights()
    with torch.no_grad():
        for s, long in FALCON_MAP.items():
            m = layer
            for p in long.split("."):
                m = getattr(m, p)
            m.weight.copy_(W[s])
        layer.ln_attn.weight.copy_(W["ln1"]); layer.ln_attn.bias.copy_(W["ln1.bias"])
        layer.ln_mlp.weight.copy_(W["ln2"]); layer.ln_mlp.bias.copy_(W["ln2.bias"])
    qlayer = QuantFalconDecoderLayer(cfg, layer.half().cuda(), args).cuda()
    x = fx.inps.cuda().to(torch.bfloat16)
    T = fx.T
    qlayer.set_quant_state(weight_quant=False, act_quant=False)
    with torch.no_grad():
        fp = torch.stack([qlayer(x[j:j + 1])[0][0] for j in range(fx.nsamples)])
    assert float((fp.float().cpu() - fx.fp_out).norm() / fx.fp_out.norm()) < 1e-2
    qlayer.set_quant_state(weight_quant=False, act_quant=True)
    qlayer.let = False
    qlayer.float()
    assert len(list(qlayer.let_parameters(True))) == 0 and len(list(qlayer.lwc_parameters())) == 8
    opt = torch.optim.AdamW([{"params": qlayer.lwc_parameters(), "lr": 1e-2}], weight_decay=0)
    loss_fn = torch.nn.MSELoss()
    losses = []
    for _ in range(fx.epochs):
        for j in range(fx.nsamples):
            qlayer.smooth_and_quant_temporary(seq_len=T, batch=1)
            out = qlayer(x[j:j + 1])[0]
            loss = loss_fn(fp[j:j + 1].float(), out.float())
            losses.append(float(loss))
            opt.zero_grad()
            loss.backward()
            opt.step()
    ref = fx.losses
    rel = np.abs(np.array(losses) - ref) / ref
    print("falcon module-loop losses", losses[:4], "ref", ref[:4], "max rel", rel.max())
    assert rel.max() < 8e-2 and abs(np.mean(losses) - ref.mean()) / ref.mean() < 3e-2
    qlayer.clear_temp_variable()
    qlayer.smooth_and_quant_inplace()
    qlayer.register_scales_and_zeros()
    lin = {n: m for n, m in qlayer.named_modules() if isinstance(m, QuantLinear)}
    assert sorted(lin) == sorted(FALCON_MAP.values())
    Q = fx.quantized()
    for s, long in FALCON_MAP.items():
        w = lin[long].weight.float().cpu()
        assert w.shape == Q[f"{s}.w"].shape
        assert float((w - Q[f"{s}.w"]).abs().mean() / Q[f"{s}.w"].abs().mean()) < 5e-2
        assert lin[long].weight_quantizer.scales.shape == Q[f"{s}.scales"].shape
        assert all(len(torch.unique(row)) <= 2 ** fx.wbits for row in w.reshape(-1, fx.group_size)[:16])
    qlayer.let = True
    with pytest.raises(ValueError):
        qlayer.smooth_and_quant_temporary(seq_len=T, batch=1)


class _Holder:
    """A bare module-like object carrying weight / bias the way QuantLinear / Omni*Norm do for the LET folds."""

    def __init__(self, weight, bias=None):
        self.weight, self.bias = weight, bias
        self._buffers = {}

    def register_buffer(self, name, value):
        setattr(self, name, value)


@pytest.mark.parametrize("fname", ["llama_w4a4_lwc_let.npz", "opt_w4a4_lwc_let.npz"])
def test_smooth_functions_match_the_reference_folds(fname):
    """The six LET fold functions (models/transformation.py:24-114) on the CUDA fold kernel, against the oracle's
    restatement of the same algebra evaluated in fp32 on the CPU at the fixture's LEARNED LET parameters: folded weights
    bit for bit (every element is one or two individually rounded fp32 operations), biases to summation-order round-off;
    then the folded weights are quantised and compared with the reference's own in-place result (fixture qw/*), and the
    gradients of the temporary variant w.r.t. scales / shifts are checked against torch autograd of the oracle."""
    from omniquant_b200 import ops
    from omniquant_b200.models import transformation as TR
    from oracle import oq_oracle as orc
    fx = BlockFixture(fname)
    W, P = fx.weights(), fx.final_params()
    llama = fx.family == "llama"
    fc1s = ["up", "gate"] if llama else ["fc1"]
    cu = lambda t: None if t is None else t.clone().cuda()
    bias = lambda n: W.get(f"{n}.bias")

    def modules():
        lin = {n: _Holder(cu(W[n]), cu(bias(n))) for n in ["q", "k", "v", "o"] + fc1s}
        ln1, ln2 = _Holder(cu(W["ln1"]), cu(W.get("ln1.bias"))), _Holder(cu(W["ln2"]), cu(W.get("ln2.bias")))
        return lin, ln1, ln2

    def ref_fold(Pd):
        pair = lambda n: (W[n], bias(n))
        ln1w, ln1b, (q, k, v) = orc.fold_ln_fcs(W["ln1"], W.get("ln1.bias"), [pair("q"), pair("k"), pair("v")], Pd["qkv_s"], Pd["qkv_d"])
        ln2w, ln2b, f1 = orc.fold_ln_fcs(W["ln2"], W.get("ln2.bias"), [pair(n) for n in fc1s], Pd["fc1_s"], Pd["fc1_d"])
        v, o = orc.fold_fc_fc(v, pair("o"), Pd["out_s"], Pd["out_d"])
        q, k = orc.fold_q_k(q, k, Pd["qkt_s"])
        out = dict(q=q, k=k, v=v, o=o, ln1=(ln1w, ln1b), ln2=(ln2w, ln2b))
        out.update({n: f for n, f in zip(fc1s, f1)})
        return out

    want = ref_fold(P)
    Pc = {k: v.clone().cuda() for k, v in P.items()}

    def check(lin, ln1, ln2, wattr, battr):
        for n in ["q", "k", "v", "o"] + fc1s:
            w, b = getattr(lin[n], wattr), getattr(lin[n], battr)
            assert torch.equal(w.detach().cpu(), want[n][0]), f"{n} folded weight"
            torch.testing.assert_close(b.detach().cpu(), want[n][1], rtol=1e-5, atol=1e-6)
        for key, ln in (("ln1", ln1), ("ln2", ln2)):
            assert torch.equal(getattr(ln, wattr).detach().cpu(), want[key][0])
            torch.testing.assert_close(getattr(ln, battr).detach().cpu(), want[key][1], rtol=1e-6, atol=1e-7)

    # ---- temporary variant (differentiable), in the reference's call order (int_llama_layer.py:285-292)
    lin, ln1, ln2 = modules()
    Pg = {k: v.clone().requires_grad_(True) for k, v in Pc.items() if k.endswith(("_s", "_d"))}
    TR.smooth_ln_fcs_temporary(ln1, [lin["q"], lin["k"], lin["v"]], Pg["qkv_s"], Pg["qkv_d"])
    TR.smooth_ln_fcs_temporary(ln2, [lin[n] for n in fc1s], Pg["fc1_s"], Pg["fc1_d"])
    TR.smooth_fc_fc_temporary(lin["v"], lin["o"], Pg["out_s"], Pg["out_d"])
    TR.smooth_q_k_temporary(lin["q"], lin["k"], Pg["qkt_s"])
    check(lin, ln1, ln2, "temp_weight", "temp_bias")
    assert all(m.use_temporary_parameter for m in lin.values())
    # gradients: a fixed random linear functional of every folded tensor
    g = torch.Generator().manual_seed(7)
    coef = {n: (torch.randn(want[n][0].shape, generator=g), torch.randn(want[n][1].shape, generator=g)) for n in lin}
    loss = sum((lin[n].temp_weight * coef[n][0].cuda()).sum() + (lin[n].temp_bias * coef[n][1].cuda()).sum() for n in lin)
    loss.backward()
    Pr = {k: v.clone().requires_grad_(True) for k, v in P.items() if k.endswith(("_s", "_d"))}
    wr = ref_fold(Pr)
    sum((wr[n][0] * coef[n][0]).sum() + (wr[n][1] * coef[n][1]).sum() for n in lin).backward()
    for k in Pr:
        a, b = Pg[k].grad.cpu(), Pr[k].grad
        assert float((a - b).norm() / b.norm()) < 1e-4, (k, float((a - b).norm() / b.norm()))

    # ---- in-place variant (int_llama_layer.py:315-327), then the weight quantiser -> the reference's own buffers
    lin, ln1, ln2 = modules()
    with torch.no_grad():
        TR.smooth_ln_fcs_inplace(ln1, [lin["q"], lin["k"], lin["v"]], Pc["qkv_s"], Pc["qkv_d"])
        TR.smooth_ln_fcs_inplace(ln2, [lin[n] for n in fc1s], Pc["fc1_s"], Pc["fc1_d"])
        TR.smooth_fc_fc_inplace(lin["v"], lin["o"], Pc["out_s"], Pc["out_d"])
        TR.smooth_q_k_inplace(lin["q"], lin["k"], Pc["qkt_s"])
    check(lin, ln1, ln2, "weight", "bias")
    assert not any(m.use_temporary_parameter for m in lin.values())
    Q = fx.quantized()
    for n in ["q", "k", "v", "o"] + fc1s:
        # (i) bit for bit against the reference quantiser's arithmetic evaluated by ATen on this GPU
        kw = dict(n_bits=fx.wbits, group_size=fx.group_size or 0, up=Pc[f"{n}.up"].view(-1).contiguous(),
                  low=Pc[f"{n}.low"].view(-1).contiguous())
        out = torch.empty_like(lin[n].weight)
        ops.weight_quant_fwd(lin[n].weight.contiguous(), recip_div=True, wq_f32=out, **kw)
        want_cuda = orc.fake_quant(lin[n].weight, fx.wbits, False, fx.group_size, Pc[f"{n}.up"], Pc[f"{n}.low"], recip_div=True)
        assert torch.equal(out, want_cuda), f"{n}: quantised in-place weight differs from ATen-CUDA"
        # (ii) against the buffers the reference itself left behind (computed on the CPU: the sigmoid of a trained clipping
        # factor may differ by one ulp between the host and the device libm, which moves a scale and flips a few codes)
        ops.weight_quant_fwd(lin[n].weight.contiguous(), recip_div=False, wq_f32=out, **kw)
        frac = float((out.cpu() != Q[f"{n}.w"]).float().mean())      # measured on B200: <= 7e-3 (one scale changes a whole row)
        assert frac < 2e-2, f"{n}: {frac:.2e} of the quantised in-place weights differ from the reference's"


def test_smooth_functions_have_no_cpu_path():
    from omniquant_b200.models import transformation as TR
    from omniquant_b200._lib import OqError
    fc = _Holder(torch.randn(8, 8), None)
    ln = _Holder(torch.ones(8), None)
    with pytest.raises(OqError):
        TR.smooth_ln_fcs_temporary(ln, [fc], torch.ones(8), torch.zeros(8))


def _tiny_layer(family, H, F_, heads, kv, abits):
    """A random-init HF decoder layer of the family and its Quant* wrapper (weights raw, 8-bit or no activation quant)."""
    torch.manual_seed(0)
    fx = SimpleNamespace(wbits=4, abits=abits, group_size=None, lwc=True, let=False)
    args = make_args(fx)
    if family == "llama":
        from transformers import LlamaConfig
        from transformers.models.llama.modeling_llama import LlamaDecoderLayer
        from omniquant_b200.models.int_llama_layer import QuantLlamaDecoderLayer as Q
        cfg = LlamaConfig(hidden_size=H, intermediate_size=F_, num_attention_heads=heads, num_key_value_heads=kv,
                          num_hidden_layers=1, vocab_size=64, max_position_embeddings=2048)
        layer = LlamaDecoderLayer(cfg, 0)
    elif family == "opt":
        from transformers import OPTConfig
        from transformers.models.opt.modeling_opt import OPTDecoderLayer
        from omniquant_b200.models.int_opt_layer import QuantOPTDecoderLayer as Q
        cfg = OPTConfig(hidden_size=H, ffn_dim=F_, num_attention_heads=heads, num_hidden_layers=1, vocab_size=64,
                        max_position_embeddings=2048, word_embed_proj_dim=H)
        try:
            layer = OPTDecoderLayer(cfg, layer_idx=0)
        except TypeError:
            layer = OPTDecoderLayer(cfg)
    else:
        from transformers import FalconConfig
        from transformers.models.falcon.modeling_falcon import FalconDecoderLayer
        from omniquant_b200.models.int_falcon_layer import QuantFalconDecoderLayer as Q
        cfg = FalconConfig(hidden_size=H, num_attention_heads=heads, num_kv_heads=kv, num_hidden_layers=1, vocab_size=64,
                           new_decoder_architecture=True, bias=False, parallel_attn=True, ffn_hidden_size=F_)
        cfg._attn_implementation = "eager"
        layer = FalconDecoderLayer(cfg, 0)
    return Q(cfg, layer.half().cuda(), args).cuda()


@pytest.mark.parametrize("family,heads,kv,T", [("llama", 2, 2, 256), ("llama", 4, 2, 128), ("llama", 4, 4, 96), ("opt", 2, 2, 256),
                                               ("opt", 4, 4, 96), ("falcon", 4, 2, 256)])
@pytest.mark.parametrize("abits", [16, 8])
def test_attention_module_forward_composes_to_the_block(family, heads, kv, T, abits):
    """Quant{Llama,OPT,Falcon}Attention.forward (sub-block level, fused attention kernels when T % 128 == 0, two GEMMs and a
    softmax otherwise) composed with the norms and the MLP reproduces the block forward of the engine."""
    d = 128 if T != 96 else 64
    H, F_ = heads * d, 2 * heads * d
    q = _tiny_layer(family, H, F_, heads, kv, abits)
    q.set_quant_state(weight_quant=False, act_quant=abits < 16)
    x = (torch.randn(2, T, H, device="cuda") * 0.5).to(torch.bfloat16)
    mask = torch.triu(torch.full((T, T), torch.finfo(torch.float32).min, device="cuda"), diagonal=1)[None, None]
    pos = torch.arange(T, device="cuda")[None]
    with torch.no_grad():
        if family == "llama":
            want = q(x, attention_mask=mask, position_ids=pos)[0]
            h = x + q.self_attn(q.input_layernorm(x), attention_mask=mask, position_ids=pos)[0]
            got = h + q.mlp(q.post_attention_layernorm(h))
        elif family == "opt":
            want = q(x, attention_mask=mask)[0]
            h = x + q.self_attn(q.self_attn_layer_norm(x), attention_mask=mask)[0]
            got = h + q.fc2(torch.relu(q.fc1(q.final_layer_norm(h))))
        else:
            want = q(x)[0]
            got = x + q.self_attention(q.ln_attn(x), None, None)[0] + q.mlp(q.ln_mlp(x))
    rel = float((got.float() - want.float()).norm() / want.float().norm())
    print(family, heads, kv, T, abits, "composed vs engine block rel", rel)
    assert rel < (1e-2 if abits >= 16 else 3e-2)


def test_attention_module_backward_matches_fp32_autograd():
    """Gradients through the fused attention operator of the module path against the fp32 autograd evaluation."""
    from omniquant_b200.models._attention import causal_attention
    b, h, hkv, T, d = 2, 4, 2, 256, 128
    g = torch.Generator(device="cuda").manual_seed(3)
    mk = lambda n: (torch.randn(b, n, T, d, device="cuda", generator=g)).to(torch.bfloat16).requires_grad_(True)
    q, k, v = mk(h), mk(hkv), mk(hkv)
    out = causal_attention(q, k, v, d ** -0.5)
    dO = torch.randn(b, T, h * d, device="cuda", generator=g).to(torch.bfloat16)
    out.backward(dO)
    qf, kf, vf = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    kk, vv = kf.repeat_interleave(h // hkv, 1), vf.repeat_interleave(h // hkv, 1)
    s = (qf @ kk.transpose(2, 3)) * d ** -0.5
    s = s.masked_fill(torch.triu(torch.ones(T, T, dtype=torch.bool, device="cuda"), diagonal=1), float("-inf"))
    ref = (torch.softmax(s, -1) @ vv).transpose(1, 2).reshape(b, T, h * d)
    ref.backward(dO.float())
    rel = lambda a, r: float((a.float() - r).norm() / r.norm())
    assert rel(out, ref) < 1e-2
    assert rel(q.grad, qf.grad) < 2e-2 and rel(k.grad, kf.grad) < 2e-2 and rel(v.grad, vf.grad) < 2e-2
