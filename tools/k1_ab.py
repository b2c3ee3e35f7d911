"""A/B of the north star's K1 formulation: weight fake-quant fused into the GEMM (oq_gemm_fq) against the pre-pass the engine
uses (oq_weight_quant_fwd writing bf16 W_q once per step, then oq_gemm_bf16), LLaMA-7B forward shapes, M = 2048 tokens.
LWC-only per-channel and g128 (group extrema constant per block: the fused side only pays oq_lwc_qparams per step)."""
import sys

import torch

sys.path.insert(0, ".")
from omniquant_b200 import ops  # noqa: E402


def timed(fn, n=20, w=5):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


M = 2048
gen = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
tot = {"pre": 0.0, "gemm": 0.0, "qp": 0.0, "fused": 0.0}
for grp, bits in ((0, 4), (128, 3)):
    for name, N, K in (("qkv", 12288, 4096), ("o", 4096, 4096), ("gate|up", 22016, 4096), ("down", 4096, 11008)):
        x = torch.randn(M, K, device="cuda", generator=gen).to(torch.bfloat16)
        W = (torch.randn(N, K, device="cuda", generator=gen) * 0.02).half().float()
        G = K // grp if grp else 1
        up = torch.full((N * G,), 4.0, device="cuda")
        z = lambda *s, dt=torch.float32: torch.empty(*s, device="cuda", dtype=dt)
        wq, sc, zp, xmin, xmax = z(N, K, dt=torch.bfloat16), z(N * G), z(N * G), z(N * G), z(N * G)
        out = z(M, N, dt=torch.bfloat16)
        pre = lambda: ops.weight_quant_fwd(W, n_bits=bits, group_size=grp, up=up, low=up, wq_bf16=wq, scale_out=sc, zp_out=zp,
                                           xmin_out=xmin, xmax_out=xmax)
        pre()
        t_pre = timed(pre)
        t_gemm = timed(lambda: ops.linear_fwd(x, wq, out))
        t_qp = timed(lambda: ops.lwc_qparams(xmin, xmax, up, up, n_bits=bits, scale_out=sc, zp_out=zp))
        t_f = timed(lambda: ops.gemm_fq(x, W, sc, zp, out, n_bits=bits, group_size=grp))
        fl = 2.0 * M * N * K
        print(f"W{bits} g{grp or 'pc':>3} {name:8s} N={N:5d} K={K:5d}: pre-pass {t_pre:7.1f} us + GEMM {t_gemm:7.1f} us = {t_pre + t_gemm:7.1f} us"
              f"   |   qparams {t_qp:5.1f} us + fused GEMM {t_f:7.1f} us = {t_qp + t_f:7.1f} us   ({fl / t_f / 1e6:6.0f} TFLOP/s fused vs "
              f"{fl / t_gemm / 1e6:6.0f} plain)   fused / unfused = {(t_qp + t_f) / (t_pre + t_gemm):.2f}")
        for k, v in (("pre", t_pre), ("gemm", t_gemm), ("qp", t_qp), ("fused", t_f)):
            tot[k] += v
print(f"sum over both configs: pre-pass + GEMM {tot['pre'] + tot['gemm']:.0f} us, fused {tot['qp'] + tot['fused']:.0f} us "
      f"(ratio {(tot['qp'] + tot['fused']) / (tot['pre'] + tot['gemm']):.2f})")
