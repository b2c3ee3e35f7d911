"""Forward QuantLinear shapes of LLaMA-7B: exact-integer GEMM (oq_gemm_s8, UTCIMMA) vs the bf16 GEMM, same pipeline."""
import sys

import torch

sys.path.insert(0, ".")
from omniquant_b200 import ops  # noqa: E402

M = 2048
g = torch.Generator().manual_seed(0)


def timed(fn, n=30, w=8):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, N, K in (("qkv", 12288, 4096), ("o", 4096, 4096), ("gate|up", 22016, 4096), ("down", 4096, 11008)):
    A8 = torch.randint(-15, 16, (M, K), generator=g, dtype=torch.int8).cuda()
    B8 = torch.randint(-15, 16, (N, K), generator=g, dtype=torch.int8).cuda()
    rs, cs = torch.rand(M, device="cuda") + 0.5, torch.rand(N, device="cuda") * 0.01
    A, B = A8.to(torch.bfloat16), B8.to(torch.bfloat16)
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    t8 = timed(lambda: ops.gemm_s8(A8, B8, C, rs, cs))
    tb = timed(lambda: ops.linear_fwd(A, B, C))
    fl = 2.0 * M * N * K
    print(f"{name:8s} M={M} N={N} K={K}: s8 {t8 * 1e3:7.1f} us ({fl / t8 / 1e9:6.0f} TOP/s)   bf16 {tb * 1e3:7.1f} us ({fl / tb / 1e9:6.0f} TFLOP/s)"
          f"   speed-up {tb / t8:.2f}x", flush=True)
