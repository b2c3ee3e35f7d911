"""Time one calibration step / forward at LLaMA-7B block shapes (dev tool)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from omniquant_b200.engine import BlockEngine, QuantConfig  # noqa: E402


def main(wbits=4, abits=4, group=0, let=True):
    H, F_, heads, T = 4096, 11008, 32, 2048
    g = torch.Generator().manual_seed(0)
    r = lambda *s: (torch.randn(*s, generator=g) * 0.02).half()
    W = {"q": r(H, H), "k": r(H, H), "v": r(H, H), "o": r(H, H), "gate": r(F_, H), "up": r(F_, H), "down": r(H, F_),
         "ln1": torch.ones(H), "ln2": torch.ones(H)}
    q = QuantConfig(wbits=wbits, abits=abits, group_size=group, lwc=True, let=let)
    eng = BlockEngine("llama", W, n_heads=heads, seq_len=T, qcfg=q, eps=1e-6)
    x = torch.randn(T, H, device="cuda").to(torch.bfloat16)
    tgt = torch.randn(T, H, device="cuda").to(torch.bfloat16)

    def timed(fn, n=5, w=2):
        for _ in range(w):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n

    print("train_step ms (gpu, wall):", timed(lambda: eng.train_step(x, tgt)))
    print("prepare ms:", timed(eng.prepare))
    print("forward(quant) ms:", timed(lambda: eng.forward(x, eng.tmp, True)))
    out = eng.forward(x, eng.tmp, True)
    dout = eng.loss_and_grad(out, tgt)
    print("backward ms:", timed(lambda: eng.backward(dout)))
    print("fp forward ms:", timed(lambda: eng.forward(x, eng.fp_weights(), False)))
    print("mem GB", torch.cuda.max_memory_allocated() / 1e9)
    # CUDA graph of the whole step
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            eng.train_step(x, tgt)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(gr):
        eng.train_step(x, tgt)
    print("graph train_step ms:", timed(gr.replay))


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    main(*a[:3], bool(a[3]) if len(a) > 3 else True) if a else main()
