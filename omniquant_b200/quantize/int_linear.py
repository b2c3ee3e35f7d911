"""QuantLinear with the reference's interface (reference quantize/int_linear.py:10-69): weight / bias
buffers, ``temp_weight`` / ``temp_bias``, ``use_temporary_parameter``, ``set_quant_state``.  The
contraction runs on the tcgen05 GEMM (bf16 operands, fp32 accumulation) through the C ABI."""
import torch
import torch.nn as nn

from .. import ops
from .quantizer import UniformAffineQuantizer


class _LinearFn(torch.autograd.Function):
    """out = x @ w^T + b on oq_gemm_bf16; dgrad / wgrad read the same buffers N-major (no transposes)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x2 = x.reshape(-1, x.shape[-1]).to(torch.bfloat16).contiguous()
        wb = w.to(torch.bfloat16).contiguous()
        out = torch.empty(x2.shape[0], w.shape[0], device=x.device, dtype=torch.bfloat16)
        ops.linear_fwd(x2, wb, out, bias=None if b is None else b.float().contiguous())
        ctx.save_for_backward(x2, wb)
        ctx.meta = (x.shape, x.dtype, w.dtype, None if b is None else b.dtype)
        return out.view(*x.shape[:-1], w.shape[0]).to(x.dtype)

    @staticmethod
    def backward(ctx, gy):
        x2, wb = ctx.saved_tensors
        xs, xdt, wdt, bdt = ctx.meta
        g2 = gy.reshape(-1, gy.shape[-1]).to(torch.bfloat16).contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x2)
            ops.linear_dgrad(g2, wb, dx)
            dx = dx.view(xs).to(xdt)
        if ctx.needs_input_grad[1]:
            dw = torch.empty(wb.shape, device=wb.device, dtype=torch.float32)
            ops.linear_wgrad(g2, x2, dw)
            dw = dw.to(wdt)
        if bdt is not None and ctx.needs_input_grad[2]:
            db = torch.empty(wb.shape[0], device=wb.device)
            ops.colsum(g2, db)
            db = db.to(bdt)
        return dx, dw, db


def quant_linear(x, weight, bias):
    return _LinearFn.apply(x, weight, bias)


class QuantLinear(nn.Module):
    """Quantized Module; to activate quantization use set_quant_state (int_linear.py:10-69)."""

    def __init__(self, org_module: nn.Linear, weight_quant_params: dict = {}, act_quant_params: dict = {},
                 disable_input_quant=False):
        super().__init__()
        self.fwd_kwargs = dict()
        self.fwd_func = quant_linear
        self.register_buffer("weight", org_module.weight.data if isinstance(org_module.weight, nn.Parameter) else org_module.weight)
        if org_module.bias is not None:
            self.register_buffer("bias", org_module.bias.data if isinstance(org_module.bias, nn.Parameter) else org_module.bias)
        else:
            self.bias = None
        self.in_features = org_module.in_features
        self.out_features = org_module.out_features
        self.use_weight_quant = False
        self.use_act_quant = False
        self.weight_quantizer = UniformAffineQuantizer(**weight_quant_params, shape=org_module.weight.shape)
        if not disable_input_quant:
            self.act_quantizer = UniformAffineQuantizer(**act_quant_params)
        else:
            self.act_quantizer = None
        self.disable_input_quant = disable_input_quant
        self.use_temporary_parameter = False

    def forward(self, input: torch.Tensor):
        if self.use_temporary_parameter:
            weight = self.temp_weight
            bias = self.temp_bias
        elif self.use_weight_quant:
            weight = self.weight_quantizer(self.weight.float())
            bias = self.bias
        else:
            weight = self.weight
            bias = self.bias
        if self.use_act_quant and not self.disable_input_quant:
            input = self.act_quantizer(input)
        return self.fwd_func(input, weight, bias, **self.fwd_kwargs)

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant
