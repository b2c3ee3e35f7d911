"""Generate golden fixtures from the LIVE reference modules.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python tests/golden/make_golden.py``.  Writes ``tests/golden/*.npz``.
The fixtures pin ``oracle/oq_oracle.py`` (tests/test_oracle_golden.py) and are
the block-level parity targets of the CUDA path (tests/test_*_gpu.py).

Nothing from the reference is copied: its modules are imported, run on CPU
(fp32, the ``--deactive_amp`` arithmetic) and only inputs/outputs are stored.
The per-block loop below follows quantize/omniquant.py:157-250 step for step,
calling the reference's own block wrapper methods.
"""
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

REF = os.environ.get("OQ_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

# un-installed deps the hot path never touches
for _n in ("auto_gptq", "auto_gptq.nn_modules", "auto_gptq.nn_modules.qlinear",
           "auto_gptq.nn_modules.qlinear.qlinear_cuda"):
    sys.modules[_n] = types.ModuleType(_n)
_tc = types.ModuleType("termcolor")
_tc.colored = lambda s, *a, **k: s
sys.modules["termcolor"] = _tc
sys.path.insert(0, REF)

from quantize.quantizer import UniformAffineQuantizer  # noqa: E402
from quantize.int_linear import QuantLinear  # noqa: E402
import models.int_llama_layer as L  # noqa: E402
import models.int_opt_layer as O  # noqa: E402
import models.int_falcon_layer as FA  # noqa: E402
from transformers import LlamaConfig, OPTConfig  # noqa: E402
from transformers.models.llama.modeling_llama import LlamaDecoderLayer, rotate_half  # noqa: E402
from transformers.models.opt.modeling_opt import OPTDecoderLayer  # noqa: E402
from transformers.models.falcon.modeling_falcon import FalconConfig, FalconDecoderLayer  # noqa: E402


def f16r(t):
    """fp32 tensor holding fp16-representable values."""
    return t.half().float()


# ---------------------------------------------------------------- quantiser KATs
def quantizer_cases():
    g = torch.Generator().manual_seed(1234)
    cases = {}
    specs = [
        # name, rows, cols, n_bits, symmetric, group, lwc
        ("w4_pc_asym", 24, 96, 4, False, None, True),
        ("w3_g32_asym", 16, 128, 3, False, 32, True),
        ("w2_g64_asym", 16, 128, 2, False, 64, True),
        ("w4_g128_asym", 8, 256, 4, False, 128, True),
        ("w6_pc_asym", 12, 80, 6, False, None, True),
        ("w8_pc_asym_nolwc", 12, 80, 8, False, None, False),
        ("w4_pc_sym", 12, 96, 4, True, None, True),
        ("w4_g48_sym_deficient", 10, 100, 4, True, 48, True),
        ("w3_g128_sym", 6, 256, 3, True, 128, True),
    ]
    for name, r, c, nb, sym, grp, lwc in specs:
        W = f16r(torch.randn(r, c, generator=g) * 0.05)
        if name == "w4_pc_asym":       # an all-positive row and a constant row (scale 0 -> NaN, like the reference)
            W[0] = W[0].abs() + 0.01
            W[1] = 0.03125
        q = UniformAffineQuantizer(n_bits=nb, symmetric=sym, per_channel_axes=[0], dynamic_method="per_channel",
                                   group_size=grp, shape=W.shape, lwc=lwc)
        if lwc:
            with torch.no_grad():
                q.upbound_factor.copy_(torch.randn(q.upbound_factor.shape, generator=g) * 1.5 + 3.0)
                q.lowbound_factor.copy_(torch.randn(q.lowbound_factor.shape, generator=g) * 1.5 + 3.0)
        Wp = W.clone().requires_grad_(True)
        y = q(Wp)
        gy = torch.randn(y.shape, generator=g)
        y.backward(gy)
        s, z = q.scale.detach(), q.round_zero_point.detach()
        xg = W
        if grp:
            pad = (-c) % grp
            xg = torch.cat((W, torch.zeros(r, pad)), 1).reshape(-1, grp) if pad else W.reshape(-1, grp)
        codes = ((xg / s).round() + z).clamp(0, 2 ** nb - 1)
        d = dict(W=W, deq=y.detach(), scale=s, zp=z, codes=codes, gy=gy, dW=Wp.grad,
                 meta=np.array([nb, int(sym), grp or 0, int(lwc)]))
        if lwc:
            d.update(up=q.upbound_factor.detach(), low=q.lowbound_factor.detach(),
                     sig_up=torch.sigmoid(q.upbound_factor.detach()), sig_low=torch.sigmoid(q.lowbound_factor.detach()),
                     d_up=q.upbound_factor.grad, d_low=q.lowbound_factor.grad)
        for k, v in d.items():
            cases[f"{name}/{k}"] = v.numpy() if isinstance(v, torch.Tensor) else v
    # per-token dynamic activations, 2-D rows and 4-D [b,h,T,d]
    for name, shape, nb in [("a4_tok", (2, 24, 96), 4), ("a6_tok", (2, 24, 96), 6), ("a4_head", (2, 3, 16, 128), 4),
                            ("a8_tok", (1, 8, 64), 8)]:
        X = f16r(torch.randn(*shape, generator=g))
        if name == "a4_tok":
            X[0, 0] = X[0, 0].abs()      # all-positive token: zero point leaves [0, qmax]
            X[0, 1, 5] = 40.0            # outlier channel
        q = UniformAffineQuantizer(n_bits=nb, symmetric=False, per_channel_axes=[], dynamic_method="per_token")
        Xp = X.clone().requires_grad_(True)
        y = q(Xp)
        gy = torch.randn(y.shape, generator=g)
        y.backward(gy)
        d = dict(X=X, deq=y.detach(), scale=q.scale.detach(), zp=q.round_zero_point.detach(), gy=gy, dX=Xp.grad,
                 meta=np.array([nb, 0, 0, 0]))
        for k, v in d.items():
            cases[f"{name}/{k}"] = v.numpy() if isinstance(v, torch.Tensor) else v
    # survey KAT
    W = torch.tensor([[-1.0, -0.5, 0.0, 0.25, 0.5, 1.0, 2.0, 3.0], [0.1, -0.2, 0.3, -0.4, 0.5, -0.6, 0.7, -0.8]])
    for name, grp in (("kat_pc", None), ("kat_g4", 4)):
        q = UniformAffineQuantizer(n_bits=4, per_channel_axes=[0], dynamic_method="per_channel", group_size=grp,
                                   shape=W.shape, lwc=True)
        y = q(W)
        cases[f"{name}/W"] = W.numpy()
        cases[f"{name}/deq"] = y.detach().numpy()
        cases[f"{name}/scale"] = q.scale.detach().numpy()
        cases[f"{name}/zp"] = q.round_zero_point.detach().numpy()
        cases[f"{name}/meta"] = np.array([4, 0, grp or 0, 1])
    # p quantiser is the identity (quantizer.py:109 is checked before :111)
    q = UniformAffineQuantizer(n_bits=16, metric="fix0to1")
    x = torch.rand(4, 4)
    assert q(x) is x
    np.savez_compressed(os.path.join(OUT, "quantizer_kat.npz"), **cases)
    print("quantizer_kat.npz", len(cases), "arrays")


# ---------------------------------------------------------------- block fixtures
class LegacyRotary(torch.nn.Module):
    """transformers-4.31 style rotary module the reference expects at
    self_attn.rotary_emb (models/int_llama_layer.py:70,124)."""

    def __init__(self, dim, max_pos=2048, base=10000):
        super().__init__()
        inv = 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim))
        emb = torch.cat((torch.outer(torch.arange(max_pos).float(), inv),) * 2, dim=-1)
        self.register_buffer("cos_cached", emb.cos()[None, None], persistent=False)
        self.register_buffer("sin_cached", emb.sin()[None, None], persistent=False)

    def forward(self, x, seq_len=None):
        return self.cos_cached[:, :, :seq_len].to(x.dtype), self.sin_cached[:, :, :seq_len].to(x.dtype)


def _legacy_rope(q, k, cos, sin, position_ids):
    cos = cos.squeeze(1).squeeze(0)[position_ids].unsqueeze(1)
    sin = sin.squeeze(1).squeeze(0)[position_ids].unsqueeze(1)
    return q * cos + rotate_half(q) * sin, k * cos + rotate_half(k) * sin


L.apply_rotary_pos_emb = _legacy_rope


def _plain_linear_tensors(layer):
    """on CPU QuantLinear buffers must be plain tensors."""
    for m in layer.modules():
        if isinstance(m, torch.nn.Linear):
            w = m.weight.data
            b = m.bias.data if m.bias is not None else None
            del m.weight
            m.weight = w
            if b is not None:
                del m.bias
                m.bias = b


def make_args(wbits, abits, group, lwc, let, sym=False):
    a = SimpleNamespace(wbits=wbits, abits=abits, group_size=group, lwc=lwc, let=let, symmetric=sym)
    a.weight_quant_params = dict(n_bits=wbits, per_channel_axes=[0], symmetric=sym, dynamic_method="per_channel",
                                 group_size=group, lwc=lwc)
    tok = dict(n_bits=abits, per_channel_axes=[], symmetric=False, dynamic_method="per_token")
    a.act_quant_params = dict(tok)
    a.q_quant_params = dict(tok)
    a.k_quant_params = dict(tok)
    a.v_quant_params = dict(tok)
    a.p_quant_params = dict(n_bits=16, metric="fix0to1")
    return a


def calibrate_reference_block(qlayer, family, args, inps, act_scales, act_shifts, mask, position_ids, *, epochs,
                              batch_size, let_lr, lwc_lr, alpha, aug_loss=False):
    """quantize/omniquant.py:157-250 for one block, fp32 (--deactive_amp)."""
    nsamples = inps.shape[0]
    kw = dict(attention_mask=mask)
    if family == "llama":
        kw["position_ids"] = position_ids
    quant_inps = inps.clone()
    fp_inps = inps.clone()
    qlayer.set_quant_state(weight_quant=False, act_quant=False)
    with torch.no_grad():
        for j in range(nsamples):
            fp_inps[j] = qlayer(fp_inps[j].unsqueeze(0), **kw)[0]
    qlayer.set_quant_state(weight_quant=False, act_quant=True)
    qlayer.let = args.let
    pairs = {"q_proj": "qkv", "o_proj": "out", "up_proj": "fc1"} if family == "llama" else \
            {"q_proj": "qkv", "out_proj": "out", "fc1": "fc1"}
    if args.let:
        nq = qlayer.self_attn.q_proj.out_features
        qlayer.register_parameter("qkt_smooth_scale", torch.nn.Parameter(torch.ones(nq)))
        for name, module in qlayer.named_modules():
            if isinstance(module, QuantLinear):
                for key in pairs:
                    if key in name:
                        act = act_scales[name].clamp(min=1e-5)
                        weight = module.weight.max(dim=0)[0].clamp(min=1e-5)
                        scale = (act.pow(alpha) / weight.pow(1 - alpha)).clamp(min=1e-5)
                        shift = torch.zeros_like(scale) if family == "llama" else act_shifts[name].clone()
                        qlayer.register_parameter(f"{pairs[key]}_smooth_shift", torch.nn.Parameter(shift))
                        qlayer.register_parameter(f"{pairs[key]}_smooth_scale", torch.nn.Parameter(scale))
    init = {k: v.detach().clone() for k, v in qlayer.omni_state_dict().items()}
    with torch.no_grad():
        qlayer.float()
    opt = torch.optim.AdamW([{"params": qlayer.let_parameters(True), "lr": let_lr},
                             {"params": qlayer.lwc_parameters(), "lr": lwc_lr}], weight_decay=0)
    loss_fn = torch.nn.MSELoss()
    mask_b = mask.repeat(batch_size, 1, 1, 1)
    losses, grads0 = [], None
    for _ in range(epochs):
        for j in range(nsamples // batch_size):
            i0 = j * batch_size
            qlayer.smooth_and_quant_temporary()
            kwb = dict(kw)
            kwb["attention_mask"] = mask_b
            out = qlayer(quant_inps[i0:i0 + batch_size], **kwb)[0]
            loss = loss_fn(fp_inps[i0:i0 + batch_size], out)
            if aug_loss:
                loss = loss + loss_fn(fp_inps[i0:i0 + batch_size], out)   # block 0: fp_inps_2 == fp_inps
            losses.append(float(loss.detach()))
            opt.zero_grad()
            loss.backward()
            if grads0 is None:
                grads0 = {k: v.grad.detach().clone() for k, v in qlayer.omni_state_dict(keep_vars=True).items()
                          if v.grad is not None}
            opt.step()
    qlayer.clear_temp_variable()
    qlayer.smooth_and_quant_inplace()
    with torch.no_grad():
        for j in range(nsamples):
            quant_inps[j] = qlayer(quant_inps[j].unsqueeze(0), **kw)[0]
    qlayer.register_scales_and_zeros()
    final = {k: v.detach().clone() for k, v in qlayer.omni_state_dict().items()}
    return dict(fp_out=fp_inps, quant_out=quant_inps, losses=losses, grads0=grads0, init=init, final=final)


def _dump(path, inputs, res, qlayer, linears):
    d = {}
    for k, v in inputs.items():
        d[f"in/{k}"] = v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    d["out/fp_out"] = res["fp_out"].numpy()
    d["out/quant_out"] = res["quant_out"].numpy()
    d["out/losses"] = np.array(res["losses"], dtype=np.float64)
    for k, v in res["grads0"].items():
        d[f"grad0/{k}"] = v.numpy()
    for k, v in res["init"].items():
        d[f"init/{k}"] = v.numpy()
    for k, v in res["final"].items():
        d[f"final/{k}"] = v.numpy()
    for short, mod in linears.items():
        d[f"qw/{short}.weight"] = mod.weight.detach().numpy()
        if getattr(mod, "bias", None) is not None:
            d[f"qw/{short}.bias"] = mod.bias.detach().numpy()
        d[f"qw/{short}.scales"] = mod.weight_quantizer.scales.detach().numpy()
        d[f"qw/{short}.zeros"] = mod.weight_quantizer.zeros.detach().numpy()
    np.savez_compressed(path, **d)
    print(os.path.basename(path), f"{os.path.getsize(path)/1e6:.2f} MB", "losses", res["losses"][:3], "...",
          res["losses"][-1])


def llama_block(tag, H, F_, heads, T, nsamples, args, epochs, batch_size=1, aug_loss=False, seed=0, kv_heads=None):
    """kv_heads < heads: grouped-query attention (Llama-2-70B); the reference supports it without LET only."""
    torch.manual_seed(seed)
    kv_heads = kv_heads or heads
    cfg = LlamaConfig(hidden_size=H, intermediate_size=F_, num_attention_heads=heads, num_key_value_heads=kv_heads,
                      num_hidden_layers=1, vocab_size=64, max_position_embeddings=2048, rms_norm_eps=1e-6)
    hf = LlamaDecoderLayer(cfg, 0).half().float().eval()
    with torch.no_grad():   # non-trivial norm weights
        hf.input_layernorm.weight.copy_(f16r(1 + 0.1 * torch.randn(H)))
        hf.post_attention_layernorm.weight.copy_(f16r(1 + 0.1 * torch.randn(H)))
    hf.self_attn.rotary_emb = LegacyRotary(H // heads)
    raw = {n: p.detach().clone() for n, p in hf.named_parameters()}
    _plain_linear_tensors(hf)
    g = torch.Generator().manual_seed(1234)
    inps = f16r(torch.randn(nsamples, T, H, generator=g))
    inps[:, :, 3] *= 6.0           # an outlier channel, what LET exists for
    names = ["self_attn.q_proj", "self_attn.o_proj", "mlp.up_proj"]
    act_scales = {n: f16r(torch.randn(H, generator=g).abs() * 2 + 0.1) for n in names}
    act_scales["self_attn.q_proj"][3] *= 6.0
    mask = torch.triu(torch.full((T, T), torch.finfo(torch.float32).min), diagonal=1)[None, None]
    pos = torch.arange(T)[None]
    qlayer = L.QuantLlamaDecoderLayer(cfg, hf, args)
    res = calibrate_reference_block(qlayer, "llama", args, inps, act_scales, None, mask, pos, epochs=epochs,
                                    batch_size=batch_size, let_lr=5e-3, lwc_lr=1e-2, alpha=0.5, aug_loss=aug_loss)
    inputs = dict(inps=inps, H=H, F=F_, heads=heads, T=T, nsamples=nsamples, epochs=epochs, batch_size=batch_size,
                  wbits=args.wbits, abits=args.abits, group_size=args.group_size or 0, lwc=int(args.lwc),
                  let=int(args.let), aug_loss=int(aug_loss))
    if kv_heads != heads:
        inputs["kv_heads"] = kv_heads
    if getattr(args, "symmetric", False):
        inputs["symmetric"] = 1
    for n, p in raw.items():
        inputs[f"w/{n}"] = p.half()
    for n, s in act_scales.items():
        inputs[f"act_scale/{n}"] = s
    lin = {"q": qlayer.self_attn.q_proj, "k": qlayer.self_attn.k_proj, "v": qlayer.self_attn.v_proj,
           "o": qlayer.self_attn.o_proj, "gate": qlayer.mlp.gate_proj, "up": qlayer.mlp.up_proj,
           "down": qlayer.mlp.down_proj}
    extra = {}
    extra["ln1.weight"] = qlayer.input_layernorm.weight.detach()
    extra["ln2.weight"] = qlayer.post_attention_layernorm.weight.detach()
    if getattr(qlayer.input_layernorm, "bias", None) is not None:
        extra["ln1.bias"] = qlayer.input_layernorm.bias.detach()
        extra["ln2.bias"] = qlayer.post_attention_layernorm.bias.detach()
    for k, v in extra.items():
        res["final"][f"__buf/{k}"] = v
    _dump(os.path.join(OUT, f"llama_{tag}.npz"), inputs, res, qlayer, lin)


def opt_block(tag, H, F_, heads, T, nsamples, args, epochs, batch_size=1, seed=0):
    torch.manual_seed(seed)
    cfg = OPTConfig(hidden_size=H, ffn_dim=F_, num_attention_heads=heads, num_hidden_layers=1, vocab_size=64,
                    max_position_embeddings=2048, word_embed_proj_dim=H, dropout=0.0, attention_dropout=0.0)
    cfg._attn_implementation = "eager"
    hf = OPTDecoderLayer(cfg).half().float().eval()
    with torch.no_grad():
        for ln in (hf.self_attn_layer_norm, hf.final_layer_norm):
            ln.weight.copy_(f16r(1 + 0.1 * torch.randn(H)))
            ln.bias.copy_(f16r(0.05 * torch.randn(H)))
        for m in hf.modules():
            if isinstance(m, torch.nn.Linear):
                m.bias.copy_(f16r(0.02 * torch.randn(m.bias.shape)))
    raw = {n: p.detach().clone() for n, p in hf.named_parameters()}
    _plain_linear_tensors(hf)
    g = torch.Generator().manual_seed(1234)
    inps = f16r(torch.randn(nsamples, T, H, generator=g))
    inps[:, :, 5] *= 5.0
    names = ["self_attn.q_proj", "self_attn.out_proj", "fc1"]
    act_scales = {n: f16r(torch.randn(H, generator=g).abs() * 2 + 0.1) for n in names}
    act_shifts = {n: f16r(torch.randn(H, generator=g) * 0.1) for n in names}
    mask = torch.triu(torch.full((T, T), torch.finfo(torch.float32).min), diagonal=1)[None, None]
    qlayer = O.QuantOPTDecoderLayer(cfg, hf, args)
    res = calibrate_reference_block(qlayer, "opt", args, inps, act_scales, act_shifts, mask, None, epochs=epochs,
                                    batch_size=batch_size, let_lr=5e-3, lwc_lr=1e-2, alpha=0.5)
    inputs = dict(inps=inps, H=H, F=F_, heads=heads, T=T, nsamples=nsamples, epochs=epochs, batch_size=batch_size,
                  wbits=args.wbits, abits=args.abits, group_size=args.group_size or 0, lwc=int(args.lwc),
                  let=int(args.let), aug_loss=0)
    for n, p in raw.items():
        inputs[f"w/{n}"] = p.half()
    for n, s in act_scales.items():
        inputs[f"act_scale/{n}"] = s
    for n, s in act_shifts.items():
        inputs[f"act_shift/{n}"] = s
    lin = {"q": qlayer.self_attn.q_proj, "k": qlayer.self_attn.k_proj, "v": qlayer.self_attn.v_proj,
           "o": qlayer.self_attn.out_proj, "fc1": qlayer.fc1, "fc2": qlayer.fc2}
    res["final"]["__buf/ln1.weight"] = qlayer.self_attn_layer_norm.weight.detach()
    res["final"]["__buf/ln1.bias"] = qlayer.self_attn_layer_norm.bias.detach()
    res["final"]["__buf/ln2.weight"] = qlayer.final_layer_norm.weight.detach()
    res["final"]["__buf/ln2.bias"] = qlayer.final_layer_norm.bias.detach()
    _dump(os.path.join(OUT, f"opt_{tag}.npz"), inputs, res, qlayer, lin)


class LegacyFalconRotary(torch.nn.Module):
    """transformers-4.31 style FalconRotaryEmbedding the reference calls as maybe_rotary(q, k, past_len)
    (models/int_falcon_layer.py:52,148): q,k [batch*heads, T, d]."""

    def __init__(self, dim, base=10000):
        super().__init__()
        self.register_buffer("inv_freq", 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim)), persistent=False)

    def forward(self, q, k, past_len=0):
        T = q.shape[1]
        t = torch.arange(past_len, past_len + T).float()
        emb = torch.cat((torch.outer(t, self.inv_freq),) * 2, dim=-1)
        cos, sin = emb.cos()[None], emb.sin()[None]
        return q * cos + rotate_half(q) * sin, k * cos + rotate_half(k) * sin


def falcon_block(tag, H, heads, kvh, T, nsamples, args, epochs, seed=0):
    torch.manual_seed(seed)
    cfg = FalconConfig(hidden_size=H, num_attention_heads=heads, num_kv_heads=kvh, num_hidden_layers=1, vocab_size=64,
                       new_decoder_architecture=True, bias=False, parallel_attn=True, hidden_dropout=0.0, attention_dropout=0.0)
    cfg._attn_implementation = "eager"
    hf = FalconDecoderLayer(cfg, 0).half().float().eval()
    with torch.no_grad():
        for ln in (hf.ln_attn, hf.ln_mlp):
            ln.weight.copy_(f16r(1 + 0.1 * torch.randn(H)))
            ln.bias.copy_(f16r(0.05 * torch.randn(H)))
    hf.self_attention.maybe_rotary = LegacyFalconRotary(H // heads)
    raw = {n: p.detach().clone() for n, p in hf.named_parameters()}
    _plain_linear_tensors(hf)
    g = torch.Generator().manual_seed(1234)
    inps = f16r(torch.randn(nsamples, T, H, generator=g))
    inps[:, :, 3] *= 6.0
    mask = torch.triu(torch.ones(T, T, dtype=torch.bool), diagonal=1)[None, None]     # True = masked (:164)
    qlayer = FA.QuantFalconDecoderLayer(cfg, hf, args)
    # the per-block loop of omniquant.py for the falcon family (no position_ids, no LET)
    quant_inps, fp_inps = inps.clone(), inps.clone()
    qlayer.set_quant_state(weight_quant=False, act_quant=False)
    with torch.no_grad():
        for j in range(nsamples):
            fp_inps[j] = qlayer(fp_inps[j].unsqueeze(0), attention_mask=mask)[0]
    qlayer.set_quant_state(weight_quant=False, act_quant=True)
    qlayer.let = False
    init = {k: v.detach().clone() for k, v in qlayer.omni_state_dict().items()}
    with torch.no_grad():
        qlayer.float()
    opt = torch.optim.AdamW([{"params": qlayer.let_parameters(True), "lr": 5e-3}, {"params": qlayer.lwc_parameters(), "lr": 1e-2}],
                            weight_decay=0)
    loss_fn = torch.nn.MSELoss()
    losses, grads0 = [], None
    for _ in range(epochs):
        for j in range(nsamples):
            qlayer.smooth_and_quant_temporary()
            out = qlayer(quant_inps[j:j + 1], attention_mask=mask)[0]
            loss = loss_fn(fp_inps[j:j + 1], out)
            losses.append(float(loss.detach()))
            opt.zero_grad()
            loss.backward()
            if grads0 is None:
                grads0 = {k: v.grad.detach().clone() for k, v in qlayer.omni_state_dict(keep_vars=True).items() if v.grad is not None}
            opt.step()
    qlayer.clear_temp_variable()
    qlayer.smooth_and_quant_inplace()
    with torch.no_grad():
        for j in range(nsamples):
            quant_inps[j] = qlayer(quant_inps[j].unsqueeze(0), attention_mask=mask)[0]
    qlayer.register_scales_and_zeros()
    res = dict(fp_out=fp_inps, quant_out=quant_inps, losses=losses, grads0=grads0, init=init,
               final={k: v.detach().clone() for k, v in qlayer.omni_state_dict().items()})
    inputs = dict(inps=inps, H=H, F=4 * H, heads=heads, kv_heads=kvh, T=T, nsamples=nsamples, epochs=epochs, batch_size=1,
                  wbits=args.wbits, abits=args.abits, group_size=args.group_size or 0, lwc=1, let=0, aug_loss=0)
    for n, p in raw.items():
        inputs[f"w/{n}"] = p.half()
    lin = {"qkv": qlayer.self_attention.query_key_value, "dense": qlayer.self_attention.dense,
           "fc1": qlayer.mlp.dense_h_to_4h, "fc2": qlayer.mlp.dense_4h_to_h}
    _dump(os.path.join(OUT, f"falcon_{tag}.npz"), inputs, res, qlayer, lin)



# ---------------------------------------------------------------- activation statistics (generate_act_scale_shift.py)
def act_stats_fixture():
    """get_act_scales / get_act_shifts of the reference on a two-linear toy model; the first linear sees the
    (bf16-representable) inputs themselves, the second arbitrary fp32 activations."""
    for n in ("datasets", "datautils", "tqdm"):
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                sys.modules[n] = types.ModuleType(n)
    if not hasattr(sys.modules["datasets"], "load_dataset"):
        sys.modules["datasets"].load_dataset = None
    if not hasattr(sys.modules["datautils"], "get_loaders"):
        sys.modules["datautils"].get_loaders = None
    if not hasattr(sys.modules["tqdm"], "tqdm"):
        sys.modules["tqdm"].tqdm = lambda it, *a, **k: it
    import generate_act_scale_shift as G
    g = torch.Generator().manual_seed(77)
    nsamples, T, C, C2 = 6, 48, 200, 72
    model = torch.nn.Sequential(torch.nn.Linear(C, C2), torch.nn.Tanh(), torch.nn.Linear(C2, 16))
    xs = (torch.randn(nsamples, 1, T, C, generator=g) * (1 + 3 * torch.rand(C, generator=g)) +
          torch.randn(C, generator=g)).bfloat16().float()
    dataloader = [(xs[i],) for i in range(nsamples)]
    with torch.no_grad():
        scales = G.get_act_scales(model, dataloader, nsamples)
        shifts = G.get_act_shifts(model, dataloader, nsamples)
        h = torch.stack([torch.tanh(model[0](xs[i])) for i in range(nsamples)])
    out = {"in/x0": xs.numpy(), "in/x2": h.numpy()}
    for k, v in scales.items():
        out[f"out/scale/{k}"] = v.numpy()
    for k, v in shifts.items():
        out[f"out/shift/{k}"] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "act_stats.npz"), **out)
    print("act_stats.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    torch.set_num_threads(4)
    if sys.argv[1:] == ["act_stats"]:
        act_stats_fixture()
        sys.exit(0)
    if sys.argv[1:] == ["extra"]:    # symmetric weights (main.py --symmetric) and a W6A6 OPT block
        llama_block("sym_w4a16g64_lwc", 256, 384, 2, 32, 4, make_args(4, 16, 64, True, False, sym=True), epochs=2)
        opt_block("w6a6_lwc_let", 128, 256, 2, 32, 4, make_args(6, 6, None, True, True), epochs=2)
        sys.exit(0)
    if sys.argv[1:] == ["gqa"]:      # BASELINE config 4's quantisation (W2A16g64 LWC) on a 4-query / 2-K/V-head block
        llama_block("gqa_w2a16g64_lwc", 256, 384, 4, 32, 4, make_args(2, 16, 64, True, False), epochs=2, kv_heads=2)
        sys.exit(0)
    quantizer_cases()
    llama_block("w4a4_lwc_let", 256, 384, 2, 32, 4, make_args(4, 4, None, True, True), epochs=2)
    llama_block("w4a4_lwc_let_bs2", 256, 384, 2, 32, 4, make_args(4, 4, None, True, True), epochs=2, batch_size=2)
    llama_block("w3a16g64_lwc", 256, 384, 2, 32, 4, make_args(3, 16, 64, True, False), epochs=2)
    llama_block("w6a6_lwc_let_aug", 256, 384, 2, 32, 4, make_args(6, 6, None, True, True), epochs=1, aug_loss=True)
    llama_block("w4a16g64_lwc_let", 256, 384, 2, 32, 4, make_args(4, 16, 64, True, True), epochs=2)
    opt_block("w4a16g64_lwc", 128, 256, 2, 32, 4, make_args(4, 16, 64, True, False), epochs=2)
    opt_block("w4a16_lwc_let", 128, 256, 2, 32, 4, make_args(4, 16, None, True, True), epochs=2)
    opt_block("w4a4_lwc_let", 128, 256, 2, 32, 4, make_args(4, 4, None, True, True), epochs=2)
    falcon_block("w3a16g64_lwc", 256, 4, 2, 32, 4, make_args(3, 16, 64, True, False), epochs=2)
    llama_block("gqa_w2a16g64_lwc", 256, 384, 4, 32, 4, make_args(2, 16, 64, True, False), epochs=2, kv_heads=2)
    llama_block("sym_w4a16g64_lwc", 256, 384, 2, 32, 4, make_args(4, 16, 64, True, False, sym=True), epochs=2)
    opt_block("w6a6_lwc_let", 128, 256, 2, 32, 4, make_args(6, 6, None, True, True), epochs=2)
    act_stats_fixture()
