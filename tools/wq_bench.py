"""Weight fake-quant forward / backward kernels alone (dev tool): GB/s of algorithmic traffic per shape and group size.
Algorithmic bytes per weight: forward 4 (fp32 W) + 2 (bf16 W_q); backward 4 (W) + 2 (bf16 dW_q)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from omniquant_b200 import ops  # noqa: E402


def timed(fn, n=20, w=3):
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


def bench(N, K, bits, group, let=False):
    g = torch.Generator(device="cuda").manual_seed(0)
    W = torch.randn(N, K, generator=g, device="cuda") * 0.02
    G = K // group if group else 1
    up = torch.full((N * G,), 4.0, device="cuda")
    low = torch.full((N * G,), 4.0, device="cuda")
    wq = torch.empty(N, K, device="cuda", dtype=torch.bfloat16)
    sc, zp, mn, mx = (torch.empty(N * G, device="cuda") for _ in range(4))
    dw = (torch.randn(N, K, generator=g, device="cuda") * 1e-3).to(torch.bfloat16)
    d_up, d_low = torch.empty(N * G, device="cuda"), torch.empty(N * G, device="cuda")
    kw = {}
    if let:
        kw = dict(col_scale=torch.rand(K, device="cuda") + 0.5)
    dcol = torch.zeros(K, device="cuda") if let else None
    fwd = lambda: ops.weight_quant_fwd(W, n_bits=bits, group_size=group, up=up, low=low, wq_bf16=wq, scale_out=sc,
                                       zp_out=zp, xmin_out=mn, xmax_out=mx, **kw)
    bwd = lambda: ops.weight_quant_bwd(W, dw, mn, mx, n_bits=bits, group_size=group, up=up, low=low, d_up=d_up, d_low=d_low,
                                       scale=sc, zp=zp, dcol_acc=dcol, **kw)
    tf, tb = timed(fwd), timed(bwd)
    b = N * K * 6
    print(json.dumps(dict(N=N, K=K, bits=bits, group=group, let=let, fwd_ms=round(tf, 4), bwd_ms=round(tb, 4),
                          fwd_GBs=round(b / tf / 1e6, 1), bwd_GBs=round(b / tb / 1e6, 1))), flush=True)


if __name__ == "__main__":
    for (N, K) in ((4096, 4096), (22016, 4096), (4096, 11008)):
        for group in ((0,) if "pc" in sys.argv else (0, 64, 128)):
            bench(N, K, 4 if group == 0 else 3, group, let=(group == 0 and K == 4096))   # the K=11008 matrix (down) has no LET fold
    bench(4096, 4096, 4, 0, let=False)
    if "pc" in sys.argv:
        sys.exit(0)
    bench(28672, 8192, 2, 64)
    bench(59392, 14848, 3, 128)
    bench(14848, 59392, 3, 128)
