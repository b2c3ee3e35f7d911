"""QuantOPTDecoderLayer and QuantOPTAttention with the reference's constructor signatures and key names
(reference models/int_opt_layer.py:16-451): pre-LN OPT block, q scaled by head_dim^-0.5 then fake-quantised
per token over the full hidden dim (:96-97), K and V likewise before the head split (:124-130), ReLU MLP
(:318).  Execution is on the B200 block engine (models/_block.py, engine.py)."""
import torch
from torch import nn

from ..quantize.int_linear import QuantLinear
from ..quantize.int_matmul import QuantMatMul
from ..quantize.omni_norm import OmniLayerNorm
from ._attention import opt_attention_forward
from ._block import QuantDecoderLayerBase
from .transformation import *  # noqa: F401,F403


class QuantOPTAttention(nn.Module):
    def __init__(self, org_module: nn.Module, embed_dim: int, num_heads: int, dropout: float = 0.0,
                 is_decoder: bool = False, bias: bool = True, args=None, disable_act_quant=False):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.dropout = dropout
        self.head_dim = embed_dim // num_heads
        if self.head_dim * num_heads != self.embed_dim:
            raise ValueError(f"embed_dim must be divisible by num_heads (got `embed_dim`: {self.embed_dim}"
                             f" and `num_heads`: {num_heads}).")
        self.scaling = self.head_dim ** -0.5
        self.is_decoder = is_decoder
        mk = lambda m: QuantLinear(m, args.weight_quant_params, args.act_quant_params)
        self.k_proj = mk(org_module.k_proj)
        self.v_proj = mk(org_module.v_proj)
        self.q_proj = mk(org_module.q_proj)
        self.out_proj = mk(org_module.out_proj)
        self.qkt_matmul = QuantMatMul(args.q_quant_params, args.k_quant_params, matmul_func=torch.bmm)
        self.pv_matmul = QuantMatMul(args.p_quant_params, args.v_quant_params, matmul_func=torch.bmm)
        self.use_weight_quant = False
        self.use_act_quant = False

    def forward(self, hidden_states, key_value_states=None, past_key_value=None, attention_mask=None, layer_head_mask=None,
                output_attentions=False):
        """Causal self-attention without KV cache on the library's kernels (models/_attention.py)."""
        return opt_attention_forward(self, hidden_states, key_value_states, past_key_value, attention_mask, layer_head_mask,
                                     output_attentions)

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant
        for m in self.modules():
            if isinstance(m, (QuantLinear, QuantMatMul)):
                m.set_quant_state(weight_quant, act_quant)


class QuantOPTDecoderLayer(QuantDecoderLayerBase):
    family = "opt"
    linear_names = {"q": "self_attn.q_proj", "k": "self_attn.k_proj", "v": "self_attn.v_proj", "o": "self_attn.out_proj",
                    "fc1": "fc1", "fc2": "fc2"}
    norm_names = ("self_attn_layer_norm", "final_layer_norm")

    def __init__(self, config, ori_layer, args):
        super().__init__()
        self.embed_dim = config.hidden_size
        if not config.do_layer_norm_before:
            raise NotImplementedError("post-LN OPT (do_layer_norm_before=False, opt-350m) is not wired into the engine")
        self.self_attn = QuantOPTAttention(org_module=ori_layer.self_attn, embed_dim=self.embed_dim,
                                           num_heads=config.num_attention_heads, dropout=config.attention_dropout,
                                           is_decoder=True, bias=config.enable_bias, args=args)
        self.do_layer_norm_before = config.do_layer_norm_before
        self.dropout = config.dropout
        self.self_attn_layer_norm = OmniLayerNorm(ori_layer.self_attn_layer_norm)
        mk = lambda m: QuantLinear(m, weight_quant_params=args.weight_quant_params, act_quant_params=args.act_quant_params)
        self.fc1 = mk(ori_layer.fc1)
        self.fc2 = mk(ori_layer.fc2)
        self.final_layer_norm = OmniLayerNorm(ori_layer.final_layer_norm)
        self.type = ori_layer.fc1.weight.dtype
        self._init_common(args, config.num_attention_heads, ori_layer.self_attn_layer_norm.eps)

    def forward(self, hidden_states, attention_mask=None, layer_head_mask=None, output_attentions=False, use_cache=False,
                past_key_value=None, **kwargs):
        if past_key_value is not None or use_cache or output_attentions or layer_head_mask is not None:
            raise NotImplementedError("the calibration block runs without KV cache / head mask / attention outputs")
        return (self._run(hidden_states, attention_mask),)
