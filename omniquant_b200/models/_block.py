"""Shared implementation of the Quant*DecoderLayer block wrappers.

Keeps the reference's block-wrapper surface (reference models/int_llama_layer.py:191-368,
models/int_opt_layer.py:225-451): ``set_quant_state``, ``smooth_and_quant_temporary``,
``clear_temp_variable``, ``smooth_and_quant_inplace``, ``let_parameters`` / ``lwc_parameters`` /
``omni_parameters`` / ``omni_state_dict``, ``register_scales_and_zeros`` and a ``forward`` with the
HF decoder-layer signature -- and runs all of it on the B200 block engine (omniquant_b200/engine.py).
The sub-modules (QuantLinear, Omni*Norm, QuantMatMul) carry the reference's buffers, parameters
and checkpoint key names; the wrapper synchronises them with the engine's flat buffers.
"""
import dataclasses
import os
import weakref
from collections import OrderedDict

import torch
from torch import nn

from ..engine import BlockEngine, QuantConfig
from ..quantize.int_linear import QuantLinear
from ..quantize.int_matmul import QuantMatMul

LET_NAMES = {"qkv_s": "qkv_smooth_scale", "qkv_d": "qkv_smooth_shift", "out_s": "out_smooth_scale",
             "out_d": "out_smooth_shift", "fc1_s": "fc1_smooth_scale", "fc1_d": "fc1_smooth_shift",
             "qkt_s": "qkt_smooth_scale"}


def rope_theta_of(config) -> float:
    """Rotary base of an HF config (``rope_theta``; transformers >= 5 keeps it in ``rope_parameters``).  The reference
    deep-copies the checkpoint's rotary module (models/int_llama_layer.py:70), so a scaled variant would silently change
    the tables: anything but the default rotary type is refused here instead of being computed wrongly."""
    if config is None:
        return 10000.0
    rp = getattr(config, "rope_parameters", None) or {}
    scaling = getattr(config, "rope_scaling", None) or {}
    kind = (rp.get("rope_type") or scaling.get("rope_type") or scaling.get("type") or "default") if isinstance(rp, dict) else "default"
    if kind != "default":
        raise NotImplementedError(f"rotary scaling type {kind!r}: only the default rotary embedding is implemented")
    theta = rp.get("rope_theta") if isinstance(rp, dict) else None
    if theta is None:
        theta = getattr(config, "rope_theta", None)
    return float(theta if theta is not None else 10000.0)


class _BlockFn(torch.autograd.Function):
    """Block forward/backward on the engine, exposed to autograd so the reference's training loop
    (loss.backward(); optimizer.step()) works unchanged on the module parameters."""

    @staticmethod
    def forward(ctx, x, layer, ws, act_quant, *params):
        eng = layer._eng
        x2 = x.reshape(-1, x.shape[-1]).to(torch.bfloat16).contiguous()
        out = eng.forward(x2, ws, act_quant=act_quant).clone()
        ctx.layer, ctx.n = layer, len(params)
        ctx.trainable = ws is eng.tmp
        return out.view(x.shape).to(x.dtype)

    @staticmethod
    def backward(ctx, gout):
        layer, eng = ctx.layer, ctx.layer._eng
        if not ctx.trainable:
            raise RuntimeError("backward through a block forward that did not use smooth_and_quant_temporary()")
        g2 = gout.reshape(-1, gout.shape[-1]).to(torch.bfloat16).contiguous()
        eng.backward(g2)
        grads = layer._pull_grads()
        return (None, None, None, None) + tuple(grads)


_ENGINES = {}      # shape key -> the resident BlockEngine (see _ensure_engine)


def release_engines():
    """Drop the resident engine(s) and with them the captured CUDA graphs.  Must run before the NCCL process group is
    destroyed: tearing down a communicator while a live graph still holds its all-reduce kernel blocks at exit."""
    for e in _ENGINES.values():
        e._graph_state = None
    _ENGINES.clear()


class QuantDecoderLayerBase(nn.Module):
    family = None          # "llama" | "opt"
    linear_names = {}      # engine short name -> attribute path of the QuantLinear
    norm_names = ()        # (first norm attr, second norm attr)

    # ------------------------------------------------------------------ plumbing
    def _init_common(self, args, n_heads, eps, n_kv_heads=None, config=None):
        self._args, self._n_heads, self._eps, self._n_kv_heads = args, n_heads, eps, n_kv_heads
        self._rope_theta = rope_theta_of(config)
        self._eng = None
        self._final_ws = None
        self.use_weight_quant = False
        self.use_act_quant = False
        self.let = False

    def _mod(self, path):
        m = self
        for p in path.split("."):
            m = getattr(m, p)
        return m

    def _linears(self):
        return OrderedDict((k, self._mod(p)) for k, p in self.linear_names.items())

    def _qcfg(self):
        a = self._args
        wq = a.weight_quant_params
        return QuantConfig(wbits=wq["n_bits"], abits=a.act_quant_params["n_bits"], group_size=wq.get("group_size") or 0,
                           symmetric=wq.get("symmetric", False), lwc=wq.get("lwc", False), let=bool(self.let),
                           recip_div=getattr(a, "recip_div", True))

    def _raw_weights(self):
        W = {}
        for k, m in self._linears().items():
            W[k] = m.weight
            if getattr(m, "bias", None) is not None:
                W[f"{k}.bias"] = m.bias
        n1, n2 = (self._mod(n) for n in self.norm_names)
        W["ln1"], W["ln2"] = n1.weight, n2.weight
        if getattr(n1, "bias", None) is not None:
            W["ln1.bias"], W["ln2.bias"] = n1.bias, n2.bias
        return W

    def _ensure_engine(self, seq_len, batch, mask=None):
        e = self._eng
        q = self._qcfg()
        if e is None or e.T != seq_len or e.b != batch or e.q != q:
            dev = next(iter(self._linears().values())).weight.device
            a = self._args
            W = self._raw_weights()
            lrs = (getattr(a, "let_lr", 5e-3), getattr(a, "lwc_lr", 1e-2), getattr(a, "wd", 0.0))
            # decoder blocks of one model share shapes: the previous block's engine (device buffers + the captured CUDA
            # graph of its optimiser step) is re-loaded with this block's weights instead of being rebuilt
            key = (self.family, str(dev), seq_len, batch, dataclasses.astuple(q), self._n_heads, self._n_kv_heads, self._eps,
                   self._rope_theta, lrs, tuple(sorted((k, tuple(v.shape)) for k, v in W.items())))
            cached = _ENGINES.get(key) if (mask is None and os.environ.get("OQ_ENGINE_CACHE", "1") == "1") else None
            owner = cached._owner() if cached is not None and getattr(cached, "_owner", None) is not None else None
            if cached is not None and (owner is None or owner._eng is not cached):
                self._eng = cached.reload(W)
            else:
                self._eng = BlockEngine(self.family, W, n_heads=self._n_heads, seq_len=seq_len, qcfg=q,
                                        eps=self._eps, device=dev, batch=batch, let_lr=lrs[0], lwc_lr=lrs[1], wd=lrs[2],
                                        mask=mask, n_kv_heads=self._n_kv_heads, rope_theta=self._rope_theta)
                if mask is None:
                    _ENGINES.clear()           # one resident engine: a new shape replaces the old one
                    _ENGINES[key] = self._eng
            self._eng._owner = weakref.ref(self)
            self._push_params()
        return self._eng

    def _unit_table(self):
        """[(engine quantisation unit, engine weight group, [linear short names stacked in that unit])]."""
        if self.family == "falcon":
            return [("qkv", "qkv", ["qkv"]), ("o", "o", ["dense"]), ("fc1", "fc1", ["fc1"]), ("fc2", "fc2", ["fc2"])]
        fc1 = ["gate", "up"] if self.family == "llama" else ["fc1"]
        fc2 = "down" if self.family == "llama" else "fc2"
        return [("q", "qkv", ["q"]), ("k", "qkv", ["k"]), ("v", "qkv", ["v"]), ("o", "o", ["o"]), ("fc1", "fc1", fc1),
                ("fc2", "fc2", [fc2])]

    def _placement(self):
        """linear short name -> (unit, group, row offset in the unit, row offset in the group buffer)."""
        lin, out, goff = self._linears(), {}, {}
        for unit, group, members in self._unit_table():
            uoff = 0
            for m in members:
                out[m] = (unit, group, uoff, goff.get(group, 0))
                uoff += lin[m].out_features
                goff[group] = goff.get(group, 0) + lin[m].out_features
        return out

    def _param_map(self):
        """[(engine name, [module parameters concatenated in engine order])]."""
        out = []
        eng = self._eng
        if eng.q.let:
            for k, long in LET_NAMES.items():
                out.append((k, [getattr(self, long)]))
        if eng.q.lwc:
            lin = self._linears()
            groups = [(unit, members) for unit, _, members in self._unit_table()]
            for unit, members in groups:
                for s, attr in (("up", "upbound_factor"), ("low", "lowbound_factor")):
                    out.append((f"{unit}.{s}", [getattr(lin[m].weight_quantizer, attr) for m in members]))
        return out

    def _push_params(self):
        with torch.no_grad():
            for name, plist in self._param_map():
                dst, o = self._eng.p[name], 0
                for p in plist:
                    n = p.numel()
                    dst[o:o + n].copy_(p.detach().reshape(-1).to(dst.dtype))
                    o += n

    def _pull_params(self):
        with torch.no_grad():
            for name, plist in self._param_map():
                src, o = self._eng.p[name], 0
                for p in plist:
                    n = p.numel()
                    p.copy_(src[o:o + n].view(p.shape).to(p.dtype))
                    o += n

    def _pull_grads(self):
        grads = []
        for name, plist in self._param_map():
            src, o = self._eng.g[name], 0
            for p in plist:
                n = p.numel()
                grads.append(src[o:o + n].view(p.shape).to(p.dtype).clone())
                o += n
        return grads

    def _params_flat(self):
        return [p for _, plist in self._param_map() for p in plist]

    # ------------------------------------------------------------------ reference surface
    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        # setting weight quantization here does not affect actual forward pass
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant
        for m in self.modules():
            if isinstance(m, (QuantLinear, QuantMatMul)):
                m.set_quant_state(weight_quant, act_quant)

    def smooth_and_quant_temporary(self, seq_len=None, batch=None):
        """LET fold + weight fake-quant for the coming forward (int_llama_layer.py:279-307)."""
        e = self._eng
        T = seq_len or (e.T if e is not None else 2048)
        b = batch or (e.b if e is not None else 1)
        eng = self._ensure_engine(T, b)
        self._push_params()
        eng.prepare()
        self._publish_temp(eng)

    def _publish_temp(self, eng):
        lin = self._linears()
        t = eng.tmp
        bias_of = {"qkv": "bqkv", "o": "bo", "fc1": "bfc1", "fc2": "bfc2"}
        for k, (unit, g, _, r0) in self._placement().items():
            m = lin[k]
            m.temp_weight = eng.Wq[g][r0:r0 + m.out_features]
            bb = t[bias_of[g]]
            m.temp_bias = None if bb is None else bb[r0:r0 + m.out_features]
            m.use_temporary_parameter = True
        for n, (w, b) in zip(self.norm_names, (("ln1w", "ln1b"), ("ln2w", "ln2b"))):
            nm = self._mod(n)
            nm.temp_weight, nm.temp_bias = t[w], t[b]
            nm.use_temporary_parameter = eng.q.let

    def clear_temp_variable(self):
        for m in self._linears().values():
            for a in ("temp_weight", "temp_bias"):
                if hasattr(m, a):
                    delattr(m, a)

    @torch.no_grad()
    def smooth_and_quant_inplace(self, seq_len=None, batch=None):
        """Fold LET into the buffers and overwrite the weights by their fake-quantised values
        (int_llama_layer.py:315-332)."""
        e = self._eng
        eng = self._ensure_engine(seq_len or (e.T if e is not None else 2048), batch or (e.b if e is not None else 1))
        self._push_params()
        extra = eng.prepare(final=True)
        self._pull_params()       # truncate_number acted on the smooth scales
        lin = self._linears()
        t = eng.tmp
        bias_of = {"qkv": "bqkv", "o": "bo", "fc1": "bfc1", "fc2": "bfc2"}
        for k, (unit, g, r0, b0) in self._placement().items():
            m = lin[k]
            N = m.out_features
            bkey = bias_of[g]
            if unit in extra:
                m.weight = extra[unit]["weight"][r0:r0 + N].to(m.weight.dtype)
                if extra[unit]["scales"] is None:       # 16-bit weights: folded only, the quantiser never ran
                    m.weight_quantizer.scale = m.weight_quantizer.round_zero_point = None
                else:
                    G = extra[unit]["scales"].numel() // extra[unit]["weight"].shape[0]
                    m.weight_quantizer.scale = extra[unit]["scales"][r0 * G:(r0 + N) * G].view(-1, 1)
                    m.weight_quantizer.round_zero_point = extra[unit]["zeros"][r0 * G:(r0 + N) * G].view(-1, 1)
            bb = t[bkey]
            if bb is not None:
                val = bb[b0:b0 + N].clone()
                if getattr(m, "bias", None) is not None:
                    m.bias = val.to(m.bias.dtype)
                else:
                    if hasattr(m, "bias"):
                        del m.bias
                    m.register_buffer("bias", val)
            m.use_temporary_parameter = False
        for n, (w, b) in zip(self.norm_names, (("ln1w", "ln1b"), ("ln2w", "ln2b"))):
            nm = self._mod(n)
            nm.weight = t[w].clone().to(nm.weight.dtype)
            if t[b] is not None:
                if getattr(nm, "bias", None) is not None:
                    nm.bias = t[b].clone().to(nm.bias.dtype)
                else:
                    if hasattr(nm, "bias"):
                        del nm.bias
                    nm.register_buffer("bias", t[b].clone())
            nm.use_temporary_parameter = False
        # operands of the post-calibration forward (raw-weight branch of QuantLinear.forward, int_linear.py:56-58)
        self._final_ws = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in t.items() if k not in ("W", "W8")}
        self._final_ws["W"] = {g: w.clone() for g, w in eng.Wq.items()}
        self._final_ws["W8"] = None if eng.W8 is None else {g: o.clone() for g, o in eng.W8.items()}

    def _named_omni(self, use_shift=True):
        template = "smooth" if use_shift else "smooth_scale"
        for n, p in self.named_parameters():
            yield n, p, (n.find("bound_factor") > -1), (n.find(template) > -1)

    def let_parameters(self, use_shift=True):
        return iter([p for _, p, _, is_let in self._named_omni(use_shift) if is_let])

    def lwc_parameters(self):
        return iter([p for _, p, is_lwc, _ in self._named_omni() if is_lwc])

    def omni_parameters(self, use_shift=True):
        return iter([p for _, p, is_lwc, is_let in self._named_omni(use_shift) if is_lwc or is_let])

    def omni_state_dict(self, destination=None, prefix="", keep_vars=False):
        if destination is None:
            destination = OrderedDict()
        for name, param in self.named_parameters():
            if name.find("smooth") > -1 or name.find("bound_factor") > -1:
                destination[prefix + name] = param if keep_vars else param.detach()
        return destination

    def register_scales_and_zeros(self):
        for m in self._linears().values():
            m.weight_quantizer.register_scales_and_zeros()

    # ------------------------------------------------------------------ forward
    def _run(self, hidden_states, attention_mask):
        b, T, H = hidden_states.shape
        mask = None
        if attention_mask is not None:
            am = attention_mask.reshape(-1, T, T)[0]
            causal = torch.triu(torch.full((T, T), torch.finfo(torch.float32).min, device=am.device), diagonal=1)
            is_causal = bool(((am.float() < -1e4) == (causal < -1e4)).all())
            if not is_causal:
                mask = attention_mask.float()
        eng = self._ensure_engine(T, b, mask=mask)
        lin = next(iter(self._linears().values()))
        if lin.use_temporary_parameter:
            ws = eng.tmp
        elif self._final_ws is not None:
            ws = self._final_ws
        elif self.use_weight_quant:
            self._push_params()
            eng.prepare()
            ws = eng.tmp
        else:
            ws = eng.fp_weights()
        return _BlockFn.apply(hidden_states, self, ws, bool(self.use_act_quant), *self._params_flat())
