"""QuantLinear fwd+bwd GEMM microbenchmark: 6*M*N*K flops per (M,K,N)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from omniquant_b200 import ops  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    res = []
    for M in (2048, 4096, 16384):
        for K, N in ((4096, 4096), (4096, 11008), (11008, 4096)):
            x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
            w = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
            dy = torch.randn(M, N, device="cuda").to(torch.bfloat16)
            y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            dx = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
            dw = torch.empty(N, K, device="cuda", dtype=torch.bfloat16)
            fl = 2.0 * M * N * K
            t_f = timeit(lambda: ops.linear_fwd(x, w, y))
            t_d = timeit(lambda: ops.linear_dgrad(dy, w, dx))
            t_w = timeit(lambda: ops.linear_wgrad(dy, x, dw))
            t_ref = timeit(lambda: torch.matmul(x, w.t()))
            r = dict(M=M, K=K, N=N, fwd_tflops=fl / t_f / 1e9, dgrad_tflops=fl / t_d / 1e9, wgrad_tflops=fl / t_w / 1e9,
                     fwd_bwd_tflops=3 * fl / (t_f + t_d + t_w) / 1e9, cublas_fwd_tflops=fl / t_ref / 1e9)
            print(json.dumps(r), flush=True)
            res.append(r)
    return res


if __name__ == "__main__":
    main()
