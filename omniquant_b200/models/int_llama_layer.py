"""QuantLlamaDecoderLayer and its parts with the reference's constructor signatures, attribute and
checkpoint key names (reference models/int_llama_layer.py:20-368).  The sub-modules hold the
buffers / parameters; the block forward, its backward and the step preparation run on the B200
block engine (see models/_block.py, engine.py) -- RoPE (transformers-4.31 rotate_half convention,
:124-125), per-(token, head) Q/K/V fake-quant (:141-142,:162), fp32 softmax (:160), SwiGLU (:44-45)."""
import torch
from torch import nn

from ..quantize.int_linear import QuantLinear
from ..quantize.int_matmul import QuantMatMul
from ..quantize.omni_norm import OmniLlamaRMSNorm
from ._attention import llama_attention_forward
from ._block import QuantDecoderLayerBase, rope_theta_of
from .transformation import *  # noqa: F401,F403  (the reference re-exports the folds from here)


class QuantLlamaMLP(nn.Module):
    def __init__(self, org_module: nn.Module, hidden_size: int, intermediate_size: int, hidden_act: str, args=None):
        super().__init__()
        mk = lambda m: QuantLinear(m, args.weight_quant_params, args.act_quant_params)
        self.gate_proj = mk(org_module.gate_proj)
        self.down_proj = mk(org_module.down_proj)
        self.up_proj = mk(org_module.up_proj)
        if hidden_act not in ("silu", "swish"):
            raise NotImplementedError(f"hidden_act {hidden_act}: the fused SwiGLU kernel implements silu")
        self.hidden_act = hidden_act

    def forward(self, x):
        h = torch.cat([self.gate_proj(x), self.up_proj(x)], dim=-1)
        g, u = h.chunk(2, dim=-1)
        return self.down_proj(torch.nn.functional.silu(g) * u)


class QuantLlamaAttention(nn.Module):
    """The attention projections and the two QuantMatMul operators (int_llama_layer.py:48-187); ``forward`` for a causal
    sequence without KV cache on the fused attention kernels (models/_attention.py)."""

    def __init__(self, org_module: nn.Module, config, args=None):
        super().__init__()
        self.config = config
        self.hidden_size = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.head_dim = self.hidden_size // self.num_heads
        self.num_key_value_heads = getattr(config, "num_key_value_heads", self.num_heads)
        self.num_key_value_groups = self.num_heads // self.num_key_value_heads
        self.max_position_embeddings = getattr(config, "max_position_embeddings", 2048)
        if self.head_dim * self.num_heads != self.hidden_size:
            raise ValueError(f"hidden_size must be divisible by num_heads (got `hidden_size`: {self.hidden_size}"
                             f" and `num_heads`: {self.num_heads}).")
        mk = lambda m: QuantLinear(m, args.weight_quant_params, args.act_quant_params)
        self.k_proj = mk(org_module.k_proj)
        self.v_proj = mk(org_module.v_proj)
        self.q_proj = mk(org_module.q_proj)
        self.o_proj = mk(org_module.o_proj)
        self.qkt_matmul = QuantMatMul(args.q_quant_params, args.k_quant_params, matmul_func=torch.matmul)
        self.pv_matmul = QuantMatMul(args.p_quant_params, args.v_quant_params, matmul_func=torch.matmul)
        self.use_weight_quant = False
        self.use_act_quant = False
        self.rope_theta = rope_theta_of(config)

    def forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value=None, output_attentions=False,
                use_cache=False):
        return llama_attention_forward(self, hidden_states, attention_mask, position_ids, past_key_value, output_attentions,
                                       use_cache, rope_theta=self.rope_theta)

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant
        for m in self.modules():
            if isinstance(m, (QuantLinear, QuantMatMul)):
                m.set_quant_state(weight_quant, act_quant)


class QuantLlamaDecoderLayer(QuantDecoderLayerBase):
    family = "llama"
    linear_names = {"q": "self_attn.q_proj", "k": "self_attn.k_proj", "v": "self_attn.v_proj", "o": "self_attn.o_proj",
                    "gate": "mlp.gate_proj", "up": "mlp.up_proj", "down": "mlp.down_proj"}
    norm_names = ("input_layernorm", "post_attention_layernorm")

    def __init__(self, config, ori_layer, args):
        super().__init__()
        self.hidden_size = config.hidden_size
        self.self_attn = QuantLlamaAttention(org_module=ori_layer.self_attn, config=config, args=args)
        self.mlp = QuantLlamaMLP(org_module=ori_layer.mlp, hidden_size=self.hidden_size,
                                 intermediate_size=config.intermediate_size, hidden_act=config.hidden_act, args=args)
        eps = getattr(ori_layer.input_layernorm, "variance_epsilon", getattr(config, "rms_norm_eps", 1e-6))
        self.input_layernorm = OmniLlamaRMSNorm(ori_layer.input_layernorm, eps=eps)
        self.post_attention_layernorm = OmniLlamaRMSNorm(ori_layer.post_attention_layernorm, eps=eps)
        self._init_common(args, config.num_attention_heads, eps,
                          getattr(config, "num_key_value_heads", config.num_attention_heads), config=config)

    def forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value=None,
                output_attentions=False, use_cache=False, **kwargs):
        if past_key_value is not None or use_cache or output_attentions:
            raise NotImplementedError("the calibration block runs without KV cache / attention outputs")
        if position_ids is not None:
            T = hidden_states.shape[1]
            if not torch.equal(position_ids.reshape(-1)[:T].cpu(), torch.arange(T)):
                raise NotImplementedError("position_ids other than arange(seq_len)")
        return (self._run(hidden_states, attention_mask),)
