"""QuantFalconDecoderLayer and its parts with the reference's constructor signatures and key names (reference
models/int_falcon_layer.py:20-397), for the new decoder architecture (Falcon-40B / 180B): two LayerNorms on the block
input, attention and MLP in parallel, fused QKV laid out per K/V group as [query heads | k | v] (:76-86), rotary on
q and k (:148), plain softmax attention without Q/K/V quantisation (:170-176), exact GELU (:25).  LWC only -- the
reference raises on LET for Falcon (:329-331,:344-346).  Execution is on the B200 block engine."""
import torch
from torch import nn

from ..quantize.int_linear import QuantLinear
from ..quantize.omni_norm import OmniLayerNorm
from ._attention import falcon_attention_forward
from ._block import QuantDecoderLayerBase, rope_theta_of


class QuantFalconMLP(nn.Module):
    def __init__(self, org_module: nn.Module, args=None):
        super().__init__()
        self.dense_h_to_4h = QuantLinear(org_module.dense_h_to_4h, args.weight_quant_params, args.act_quant_params)
        self.act = nn.GELU()
        self.dense_4h_to_h = QuantLinear(org_module.dense_4h_to_h, args.weight_quant_params, args.act_quant_params)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.dense_4h_to_h(self.act(self.dense_h_to_4h(x)))


class QuantFalconAttention(nn.Module):
    """Holder of the fused QKV and output projections (int_falcon_layer.py:35-62)."""

    def __init__(self, config, org_module: nn.Module, args=None):
        super().__init__()
        self.config = config
        self.hidden_size = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.head_dim = self.hidden_size // self.num_heads
        if self.head_dim * self.num_heads != self.hidden_size:
            raise ValueError(f"`hidden_size` must be divisible by num_heads (got `hidden_size`: {self.hidden_size} and "
                             f"`num_heads`: {self.num_heads}).")
        self.new_decoder_architecture = config.new_decoder_architecture
        self.multi_query = config.multi_query
        self.num_kv_heads = config.num_kv_heads if (self.new_decoder_architecture or not self.multi_query) else 1
        self.inv_norm_factor = 1.0 / self.head_dim ** 0.5
        self.query_key_value = QuantLinear(org_module.query_key_value, args.weight_quant_params, args.act_quant_params)
        self.dense = QuantLinear(org_module.dense, args.weight_quant_params, args.act_quant_params)
        self.rope_theta = rope_theta_of(config)

    def forward(self, hidden_states, alibi, attention_mask, layer_past=None, head_mask=None, use_cache=False,
                output_attentions=False):
        """Causal rotary attention without KV cache on the fused attention kernels (models/_attention.py)."""
        return falcon_attention_forward(self, hidden_states, alibi, attention_mask, layer_past, head_mask, use_cache,
                                        output_attentions, rope_theta=self.rope_theta)


class QuantFalconDecoderLayer(QuantDecoderLayerBase):
    family = "falcon"
    linear_names = {"qkv": "self_attention.query_key_value", "dense": "self_attention.dense",
                    "fc1": "mlp.dense_h_to_4h", "fc2": "mlp.dense_4h_to_h"}
    norm_names = ("ln_attn", "ln_mlp")

    def __init__(self, config, ori_layer, args):
        super().__init__()
        if not config.new_decoder_architecture:
            raise NotImplementedError("Falcon blocks are implemented for new_decoder_architecture=True (Falcon-40B/180B); "
                                      "the 7B layout (single LayerNorm, multi-query) is not wired into the engine")
        if getattr(config, "alibi", False):
            raise NotImplementedError("alibi attention bias")
        self.num_heads = config.num_attention_heads
        self.config = config
        self.self_attention = QuantFalconAttention(config, ori_layer.self_attention, args)
        self.mlp = QuantFalconMLP(ori_layer.mlp, args)
        self.hidden_dropout = config.hidden_dropout
        self.ln_attn = OmniLayerNorm(ori_layer.ln_attn)
        self.ln_mlp = OmniLayerNorm(ori_layer.ln_mlp)
        self._init_common(args, config.num_attention_heads, ori_layer.ln_attn.eps, config.num_kv_heads, config=config)

    def smooth_and_quant_temporary(self, *a, **k):
        if self.let:
            raise ValueError("falcon not yet support let")
        return super().smooth_and_quant_temporary(*a, **k)

    def smooth_and_quant_inplace(self, *a, **k):
        if self.let:
            raise ValueError("falcon not yet support let")
        return super().smooth_and_quant_inplace(*a, **k)

    def forward(self, hidden_states, attention_mask=None, alibi=None, layer_past=None, head_mask=None, use_cache=False,
                output_attentions=False, **kwargs):
        if alibi is not None or layer_past is not None or head_mask is not None or use_cache or output_attentions:
            raise NotImplementedError("the calibration block runs without alibi / KV cache / head mask / attention outputs")
        # the reference takes a BOOL mask (True = masked, int_falcon_layer.py:164); calibration uses the causal one
        if attention_mask is not None and attention_mask.dtype == torch.bool:
            T = hidden_states.shape[1]
            causal = torch.triu(torch.ones(T, T, dtype=torch.bool, device=attention_mask.device), diagonal=1)
            if not torch.equal(attention_mask.reshape(-1, T, T)[0], causal):
                raise NotImplementedError("non-causal boolean attention masks")
            attention_mask = None
        return (self._run(hidden_states, attention_mask),)
