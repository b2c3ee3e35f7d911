"""UniformAffineQuantizer with the reference's constructor, attributes and semantics
(reference quantize/quantizer.py:23-153), computing on the B200 through the C-ABI kernels.

fp32 2-D inputs (weights; per-channel / per-group, LWC) go through ``oq_weight_quant_fwd/bwd``
and give bit-exact codes / scales / zero points; bf16 inputs (activations, per-token over the last
dim) go through ``oq_act_quant_fwd/bwd``.  There is no CPU path: a CPU tensor raises.
"""
import math

import torch
import torch.nn as nn

from .. import ops
from .._lib import OqError

CLIPMIN = 1e-5


class _WeightFakeQuant(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, up, low, qz):
        N, K = x.shape
        G = math.ceil(K / qz.group_size) if qz.group_size else 1
        dev = x.device
        out = torch.empty_like(x)
        scale, zp = torch.empty(N * G, 1, device=dev), torch.empty(N * G, 1, device=dev)
        xmin, xmax = torch.empty(N * G, device=dev), torch.empty(N * G, device=dev)
        ops.weight_quant_fwd(x.contiguous(), n_bits=qz.n_bits, symmetric=qz.symmetric, group_size=qz.group_size or 0,
                             up=up, low=low, recip_div=qz.recip_div, wq_f32=out, scale_out=scale, zp_out=zp,
                             xmin_out=xmin, xmax_out=xmax)
        qz.scale, qz.round_zero_point = scale, zp
        ctx.save_for_backward(x, up, low, xmin, xmax)
        ctx.qz = qz
        return out

    @staticmethod
    def backward(ctx, gy):
        x, up, low, xmin, xmax = ctx.saved_tensors
        qz = ctx.qz
        dW = torch.empty_like(x)
        d_up = torch.empty_like(up) if up is not None else None
        d_low = torch.empty_like(low) if low is not None else None
        ops.weight_quant_bwd(x.contiguous(), gy.contiguous().float(), xmin, xmax, n_bits=qz.n_bits,
                             symmetric=qz.symmetric, group_size=qz.group_size or 0, up=up, low=low,
                             recip_div=qz.recip_div, d_up=d_up, d_low=d_low, dW_full=dW)
        return dW, d_up, d_low, None


class _ActFakeQuant(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, qz):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        y = torch.empty_like(x2)
        rows = x2.shape[0]
        scale, zp = torch.empty(rows, device=x.device), torch.empty(rows, device=x.device)
        ops.act_quant_fwd(x2, y, n_bits=qz.n_bits, symmetric=qz.symmetric, recip_div=qz.recip_div, scale_out=scale,
                          zp_out=zp)
        shp = tuple(x.shape[:-1]) + (1,)
        qz.scale, qz.round_zero_point = scale.view(shp), zp.view(shp)
        ctx.save_for_backward(x2)
        ctx.qz, ctx.shape = qz, x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, gy):
        (x2,) = ctx.saved_tensors
        qz = ctx.qz
        g2 = gy.reshape(x2.shape).contiguous().to(torch.bfloat16)
        gx = torch.empty_like(x2)
        ops.act_quant_bwd(x2, g2, gx, n_bits=qz.n_bits, symmetric=qz.symmetric, recip_div=qz.recip_div)
        return gx.view(ctx.shape), None


class UniformAffineQuantizer(nn.Module):
    def __init__(self, n_bits: int = 8, symmetric: bool = False, per_channel_axes=[], metric="minmax", dynamic=False,
                 dynamic_method="per_cluster", group_size=None, shape=None, lwc=False):
        super().__init__()
        self.symmetric = symmetric
        assert 2 <= n_bits <= 16, "bitwidth not supported"
        self.n_bits = n_bits
        self.qmin = 0
        self.qmax = 2 ** (n_bits) - 1
        self.per_channel_axes = per_channel_axes
        self.metric = metric
        self.cluster_counts = None
        self.cluster_dim = None
        self.scale = None
        self.zero_point = None
        self.round_zero_point = None
        self.cached_xmin = None
        self.cached_xmax = None
        self.dynamic = dynamic
        self.dynamic_method = dynamic_method
        self.deficiency = 0
        self.lwc = lwc
        self.recip_div = True      # ATen-CUDA semantics of `range / (2**n - 1)` (the reference's GPU path)
        init_value = 4.0            # quantizer.py:62
        if lwc:
            if group_size:
                dim1 = int(shape[0] * math.ceil(shape[1] / group_size))
                self.deficiency = shape[-1] % group_size
                if self.deficiency > 0:
                    self.deficiency = group_size - self.deficiency
                    assert self.symmetric   # quantizer.py:69
            else:
                dim1 = shape[0]
            self.upbound_factor = nn.Parameter(torch.ones((dim1, 1)) * init_value)
            self.lowbound_factor = nn.Parameter(torch.ones((dim1, 1)) * init_value)
        self.sigmoid = nn.Sigmoid()
        self.enable = True
        self.group_size = group_size

    def change_n_bits(self, n_bits):
        self.n_bits = n_bits
        self.qmin = 0
        self.qmax = 2 ** (n_bits) - 1

    def forward(self, x: torch.Tensor):
        if self.n_bits >= 16 or not self.enable:          # quantizer.py:109 (checked before fix0to1)
            return x
        if self.metric == "fix0to1":
            return x.mul_(2 ** self.n_bits - 1).round_().div_(2 ** self.n_bits - 1)
        if self.dynamic_method not in ("per_token", "per_channel"):
            raise NotImplementedError()
        if not x.is_cuda:
            raise OqError("UniformAffineQuantizer: CUDA tensors only (no CPU fallback)")
        if self.n_bits > 8:
            raise OqError("UniformAffineQuantizer: the CUDA kernels cover 2..8 bits (and >=16 = identity)")
        if x.dtype == torch.float32 and x.dim() == 2 and (self.group_size or self.lwc or self.dynamic_method == "per_channel"):
            up = self.upbound_factor if self.lwc else None
            low = self.lowbound_factor if self.lwc else None
            return _WeightFakeQuant.apply(x, up, low, self)
        if self.group_size:
            raise OqError("grouped quantisation is implemented for fp32 2-D weights")
        return _ActFakeQuant.apply(x.to(torch.bfloat16), self).to(x.dtype)

    def register_scales_and_zeros(self):
        self.register_buffer("scales", self.scale)
        self.register_buffer("zeros", self.round_zero_point)
        del self.scale
        del self.round_zero_point
