"""``--real_quant``: the calibrated QuantLinear modules packed into the GPTQ wire format the reference hands to
AutoGPTQ's ``qlinear_cuda.QuantLinear`` (reference quantize/omniquant.py:255-278).  The packing runs on the GPU
(oq_gptq_pack); the container module keeps AutoGPTQ's buffer names so ``save_pretrained`` writes the same keys
(qweight / qzeros / scales / g_idx / bias).  Inference on the packed format is out of scope of the calibration path."""
import torch
from torch import nn

from .. import ops
from .int_linear import QuantLinear


class PackedQuantLinear(nn.Module):
    """Buffers of auto_gptq.nn_modules.qlinear.qlinear_cuda.QuantLinear(bits, group_size, infeatures, outfeatures, bias)."""

    def __init__(self, bits, group_size, infeatures, outfeatures, bias):
        super().__init__()
        if bits not in (2, 3, 4, 8):
            raise NotImplementedError("Only 2,3,4,8 bits are supported.")
        self.bits, self.infeatures, self.outfeatures = bits, infeatures, outfeatures
        self.group_size = group_size if group_size not in (None, 0, -1) else infeatures
        G = infeatures // self.group_size
        self.register_buffer("qweight", torch.zeros((infeatures // 32 * bits, outfeatures), dtype=torch.int32))
        self.register_buffer("qzeros", torch.zeros((G, outfeatures // 32 * bits), dtype=torch.int32))
        self.register_buffer("scales", torch.zeros((G, outfeatures), dtype=torch.float16))
        self.register_buffer("g_idx", torch.arange(infeatures, dtype=torch.int32) // self.group_size)
        if bias:
            self.register_buffer("bias", torch.zeros(outfeatures, dtype=torch.float16))
        else:
            self.bias = None

    def pack(self, linear, scales, zeros, g_idx=None):
        """Same call as AutoGPTQ's: ``linear.weight`` fake-quantised, scales / zeros [out, n_groups]."""
        if g_idx is not None:
            raise NotImplementedError("act-order (custom g_idx) is not produced by the calibration path")
        W = linear.weight.detach().float().contiguous()
        if not W.is_cuda:
            raise RuntimeError("PackedQuantLinear.pack runs on the GPU: move the module to the device first")
        qw, qz, sc, gi = ops.gptq_pack(W, scales.detach().float().contiguous().to(W.device),
                                       zeros.detach().float().contiguous().to(W.device), n_bits=self.bits,
                                       group_size=0 if self.group_size == self.infeatures else self.group_size)
        self.qweight, self.qzeros, self.scales, self.g_idx = qw, qz, sc, gi
        if getattr(linear, "bias", None) is not None:
            self.bias = linear.bias.detach().clone().half()

    def forward(self, x):
        raise NotImplementedError("inference kernels for the packed format are outside the calibration hot path "
                                  "(load the saved model with AutoGPTQ, as the reference does)")


def pack_block(qlayer, wbits):
    """Replace every QuantLinear of a calibrated block (scales / zeros registered) by its packed form, in place."""
    for name, module in [(n, m) for n, m in qlayer.named_modules() if isinstance(m, QuantLinear)]:
        wq = module.weight_quantizer
        dim0 = module.weight.shape[0]
        packed = PackedQuantLinear(wbits, wq.group_size, module.in_features, module.out_features, module.bias is not None)
        packed.pack(module, wq.scales.view(dim0, -1), wq.zeros.view(dim0, -1))
        parent = qlayer
        levels = name.split(".")
        for lv in levels[:-1]:
            parent = parent[int(lv)] if lv.isdigit() else getattr(parent, lv)
        setattr(parent, levels[-1], packed)
