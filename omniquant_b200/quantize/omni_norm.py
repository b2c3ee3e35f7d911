"""OmniLayerNorm / OmniLlamaRMSNorm with the reference's interface (reference
quantize/omni_norm.py:11-63): norm with swappable ``temp_weight`` / ``temp_bias``, on oq_norm_fwd/bwd."""
import torch
import torch.nn as nn

from .. import ops


class _NormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps, kind):
        x2 = x.reshape(-1, x.shape[-1]).to(torch.bfloat16).contiguous()
        y = torch.empty_like(x2)
        mean, rstd = torch.empty(x2.shape[0], device=x.device), torch.empty(x2.shape[0], device=x.device)
        wf = w.float().contiguous()
        ops.norm_fwd(x2, wf, None if b is None else b.float().contiguous(), eps, kind, y, mean, rstd)
        ctx.save_for_backward(x2, wf, mean, rstd)
        ctx.meta = (x.shape, x.dtype, kind, b is not None)
        return y.view(x.shape).to(x.dtype)

    @staticmethod
    def backward(ctx, gy):
        x2, wf, mean, rstd = ctx.saved_tensors
        shape, dt, kind, has_b = ctx.meta
        g2 = gy.reshape(x2.shape).to(torch.bfloat16).contiguous()
        dx = torch.empty_like(x2)
        dw = torch.zeros_like(wf)
        db = torch.zeros_like(wf) if has_b else None
        ops.norm_bwd(x2, wf, kind, mean, rstd, g2, dx=dx, dw_acc=dw, db_acc=db)
        return dx.view(shape).to(dt), dw, db, None, None


class OmniLayerNorm(nn.Module):
    def __init__(self, ori_layer_norm) -> None:
        super().__init__()
        self.use_act_quant = True
        self.register_buffer("weight", ori_layer_norm.weight.data)
        if ori_layer_norm.bias is not None:
            self.register_buffer("bias", ori_layer_norm.bias.data)
        else:
            self.bias = None
        self.eps = ori_layer_norm.eps
        self.normalized_shape = ori_layer_norm.normalized_shape
        self.use_temporary_parameter = False

    def forward(self, x):
        if self.use_temporary_parameter:
            weight, bias = self.temp_weight, self.temp_bias
        else:
            weight, bias = self.weight, self.bias
        return _NormFn.apply(x, weight, bias, self.eps, 1)

    def set_quant_state(self, use_weight_quant, use_act_quant):
        self.use_act_quant = use_act_quant


class OmniLlamaRMSNorm(nn.Module):
    def __init__(self, ori_norm, eps=1e-6):
        super().__init__()
        self.register_buffer("weight", ori_norm.weight.data)
        self.bias = None
        self.variance_epsilon = eps
        self.use_temporary_parameter = False

    def forward(self, hidden_states):
        if self.use_temporary_parameter:
            weight, bias = self.temp_weight, self.temp_bias
        else:
            weight = self.weight
            bias = self.bias if hasattr(self, "bias") else None
        return _NormFn.apply(hidden_states, weight, bias, self.variance_epsilon, 0)
