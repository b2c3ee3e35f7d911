"""CPU restatement of the AutoGPTQ `qlinear_cuda.QuantLinear.pack` wire format (TEST INFRASTRUCTURE ONLY -- nothing in
the product may import this file).

PARITY UNPINNED: the packer is a third-party dependency that is absent from /root/reference -- the reference calls
`auto_gptq.nn_modules.qlinear.qlinear_cuda.QuantLinear(bits, group_size, in, out, bias).pack(module, scales, zeros)`
(reference quantize/omniquant.py:255-265) from the un-pinned fork "ChenMnZ/AutoGPTQ-bugfix" (reference README.md:38-41),
and holds no golden vectors for it.  What follows restates the published upstream AutoGPTQ (v0.4.x) algorithm:

  scales^T, zeros^T           : [n_groups, out_features]
  intweight[k, n]             = round((W[n, k] + scale[g(k), n] * zero[g(k), n]) / half(scale[g(k), n])),  g(k) = k // group
  qweight  [K // 32 * bits, N]: int32; bits in {2,4,8}: 32/bits consecutive k per word, lowest k in the lowest bits;
                                bits == 3: 32 consecutive k in 3 words (10 + 1 split + 10 + 1 split + 10 codes)
  qzeros   [n_groups, N // 32 * bits]: (zero - 1) packed the same way along n  (the upstream kernels add the 1 back)
  scales   [n_groups, N] fp16,  g_idx [K] int32 = k // group

Deviation stated: every field is masked to `bits` bits before it is OR-ed into its word.  Upstream ORs the unmasked
uint32, so a zero point of 0 (field -1 = 0xFFFFFFFF) overwrites the neighbouring fields -- the kind of defect the
"bugfix" fork exists for; with the mask, zero == 0 is stored as the all-ones field and recovered by (field + 1) mod
2^bits.  In-range codes and zero points >= 1 give identical words either way.

The reference call sites pin the INPUTS: `module.weight` is the fake-quantised weight (code - zero) * scale,
scales / zeros are `weight_quantizer.scales / zeros` viewed [out, n_groups] (quantize/omniquant.py:258-263), so
intweight equals the integer code of the calibration (quantize/quantizer.py:93-96).
"""
import numpy as np


def _pack_rows(v, bits):
    """v: uint32 [R, C] -> packed along axis 0: [R // 32 * bits, C] uint32 (AutoGPTQ's loop, vectorised)."""
    R, C = v.shape
    assert R % 32 == 0
    if bits in (2, 4, 8):
        per = 32 // bits
        t = v.reshape(R // per, per, C)
        out = np.zeros((R // per, C), dtype=np.uint32)
        for j in range(per):
            out |= t[:, j] << np.uint32(bits * j)
        return out
    assert bits == 3
    t = v.reshape(R // 32, 32, C)
    out = np.zeros((R // 32, 3, C), dtype=np.uint32)
    for j in range(10):
        out[:, 0] |= t[:, j] << np.uint32(3 * j)
    out[:, 0] |= t[:, 10] << np.uint32(30)
    out[:, 1] |= (t[:, 10] >> np.uint32(2)) & np.uint32(1)
    for j in range(10):
        out[:, 1] |= t[:, 11 + j] << np.uint32(3 * j + 1)
    out[:, 1] |= t[:, 21] << np.uint32(31)
    out[:, 2] |= (t[:, 21] >> np.uint32(1)) & np.uint32(3)
    for j in range(10):
        out[:, 2] |= t[:, 22 + j] << np.uint32(3 * j + 2)
    return out.reshape(R // 32 * 3, C)


def _unpack_rows(p, bits, R):
    C = p.shape[1]
    if bits in (2, 4, 8):
        per = 32 // bits
        out = np.zeros((R // per, per, C), dtype=np.uint32)
        for j in range(per):
            out[:, j] = (p >> np.uint32(bits * j)) & np.uint32((1 << bits) - 1)
        return out.reshape(R, C)
    t = p.reshape(R // 32, 3, C)
    out = np.zeros((R // 32, 32, C), dtype=np.uint32)
    for j in range(10):
        out[:, j] = (t[:, 0] >> np.uint32(3 * j)) & np.uint32(7)
    out[:, 10] = ((t[:, 0] >> np.uint32(30)) & np.uint32(3)) | ((t[:, 1] & np.uint32(1)) << np.uint32(2))
    for j in range(10):
        out[:, 11 + j] = (t[:, 1] >> np.uint32(3 * j + 1)) & np.uint32(7)
    out[:, 21] = ((t[:, 1] >> np.uint32(31)) & np.uint32(1)) | ((t[:, 2] & np.uint32(3)) << np.uint32(1))
    for j in range(10):
        out[:, 22 + j] = (t[:, 2] >> np.uint32(3 * j + 2)) & np.uint32(7)
    return out.reshape(R, C)


def pack(W, scales, zeros, bits, group_size):
    """W [N, K] fp32 (fake-quantised), scales / zeros [N, n_groups] fp32 -> dict(qweight, qzeros, scales, g_idx)."""
    W = np.asarray(W, dtype=np.float32)
    N, K = W.shape
    group = group_size if group_size and group_size > 0 else K
    mask = (1 << bits) - 1
    g_idx = (np.arange(K) // group).astype(np.int32)
    s_t = np.ascontiguousarray(np.asarray(scales, np.float32).reshape(N, -1).T)       # [G, N]
    z_t = np.ascontiguousarray(np.asarray(zeros, np.float32).reshape(N, -1).T)
    scale_zeros = z_t * s_t
    s_half = s_t.astype(np.float16)
    num = W.T + scale_zeros[g_idx]                                                    # [K, N] fp32
    intweight = np.rint(num / s_half[g_idx].astype(np.float32)).astype(np.int32).astype(np.uint32) & np.uint32(mask)
    qweight = _pack_rows(intweight, bits).astype(np.int32)
    zm1 = (z_t - 1).astype(np.int64).astype(np.uint32) & np.uint32(mask)              # zeros -= 1, field masked (see above)
    qzeros = _pack_rows(np.ascontiguousarray(zm1.T), bits).T.copy().astype(np.int32)  # packed along n
    return dict(qweight=qweight, qzeros=qzeros, scales=s_half, g_idx=g_idx)


def unpack(packed, bits, N, K):
    """Inverse of pack: integer codes [N, K] and zero points [N, G] (int64)."""
    qw = packed["qweight"].astype(np.uint32)
    codes = _unpack_rows(qw, bits, K).T                                               # [N, K]
    qz = packed["qzeros"].astype(np.uint32)
    G = qz.shape[0]
    zm1 = _unpack_rows(np.ascontiguousarray(qz.T), bits, N).T                         # [G, N]
    zeros = (zm1.astype(np.int64) + 1) & ((1 << bits) - 1) if bits != 32 else zm1
    return codes.astype(np.int64), zeros.T.astype(np.int64)


def dequant(packed, bits, N, K):
    """What the upstream kernels compute: scale * (q - (qzero + 1)) -> [N, K] fp32."""
    codes, zeros = unpack(packed, bits, N, K)
    g = packed["g_idx"]
    s = packed["scales"].astype(np.float32).T                                         # [N, G]
    return (codes - zeros[:, g]).astype(np.float32) * s[:, g]
