"""QuantMatMul with the reference's interface (reference quantize/int_matmul.py:7-43).  ``quant_x1`` /
``quant_x2`` fake-quantise the operands per token over the last dim on the GPU kernels; ``forward``
contracts them on the tcgen05 GEMM.  Inside the calibration engine the whole QK^T -> softmax -> PV
chain is scheduled directly (omniquant_b200/engine.py); this module is the stand-alone operator."""
import torch
import torch.nn as nn

from .. import ops
from .quantizer import UniformAffineQuantizer


class _BmmFn(torch.autograd.Function):
    """x1 [..., M, K] @ x2 [..., K, N] with bf16 operands; x2 is read N-major (no transpose copy)."""

    @staticmethod
    def forward(ctx, x1, x2):
        a = x1.to(torch.bfloat16).contiguous()
        b = x2.to(torch.bfloat16).contiguous()
        M, K = a.shape[-2:]
        N = b.shape[-1]
        B = a.numel() // (M * K)
        out = torch.empty(*a.shape[:-2], M, N, device=a.device, dtype=torch.bfloat16)
        ops.gemm(a, b, out, M=M, N=N, K=K, lda=K, ldb=N, b_mn=True, ldc=N, batch=B, strideA=M * K, strideB=K * N,
                 strideC=M * N)
        ctx.save_for_backward(a, b)
        ctx.dts = (x1.dtype, x2.dtype)
        return out.to(x1.dtype)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = g.to(torch.bfloat16).contiguous()
        M, K = a.shape[-2:]
        N = b.shape[-1]
        B = a.numel() // (M * K)
        da = torch.empty_like(a)      # dA = g @ b^T : [M,N] x [K,N]^T, both K-major over N
        ops.gemm(g, b, da, M=M, N=K, K=N, lda=N, ldb=N, ldc=K, batch=B, strideA=M * N, strideB=K * N, strideC=M * K)
        db = torch.empty_like(b)      # dB = a^T @ g : [K,M] x [N,M]^T, both read MN-major over M
        ops.gemm(a, g, db, M=K, N=N, K=M, a_mn=True, lda=K, b_mn=True, ldb=N, ldc=N, batch=B, strideA=M * K,
                 strideB=M * N, strideC=K * N)
        return da.to(ctx.dts[0]), db.to(ctx.dts[1])


def quant_bmm(x1, x2):
    return _BmmFn.apply(x1, x2)


class QuantMatMul(nn.Module):
    def __init__(self, x1_quant_params: dict = {}, x2_quant_params: dict = {}, disable_act_quant=False,
                 matmul_func=torch.bmm):
        super().__init__()
        self.use_act_quant = False
        self.i_cluster_counts = None
        self.x1_quantizer = UniformAffineQuantizer(**x1_quant_params)
        self.x2_quantizer = UniformAffineQuantizer(**x2_quant_params)
        # torch.bmm / torch.matmul are accepted for signature compatibility; the contraction itself runs on
        # the sm_100a GEMM for CUDA tensors
        self.matmul_func = matmul_func
        self.disable_act_quant = disable_act_quant

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant

    def quant_x1(self, x1):
        if self.use_act_quant:
            x1 = self.x1_quantizer(x1)
        return x1

    def quant_x2(self, x2):
        if self.use_act_quant:
            x2 = self.x2_quantizer(x2)
        return x2

    def forward(self, x1, x2):
        return quant_bmm(x1, x2)
