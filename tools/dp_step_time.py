"""Cost of the per-step gradient all-reduce under data parallelism (run under torchrun): graph-replayed optimiser step with
the NCCL all-reduce captured, with the all-reduce left out (timing only -- ranks diverge), and eager with NCCL."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from omniquant_b200.engine import BlockEngine, QuantConfig  # noqa: E402
from omniquant_b200.parallel import make_allreduce  # noqa: E402

rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world = dist.get_world_size()
H, F_, heads, T = 4096, 11008, 32, 2048
g = torch.Generator().manual_seed(0)
r = lambda *s: (torch.randn(*s, generator=g) * 0.02).half()
W = {"q": r(H, H), "k": r(H, H), "v": r(H, H), "o": r(H, H), "gate": r(F_, H), "up": r(F_, H), "down": r(H, F_),
     "ln1": torch.ones(H), "ln2": torch.ones(H)}
eng = BlockEngine("llama", W, n_heads=heads, seq_len=T, qcfg=QuantConfig(wbits=4, abits=4, group_size=0, lwc=True, let=True), eps=1e-6)
eng.world = world
x = torch.randn(T, H, device="cuda").to(torch.bfloat16)
tgt = torch.randn(T, H, device="cuda").to(torch.bfloat16)
ar = make_allreduce()


def timed(fn, n=200, w=10):
    for _ in range(w):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / n * 1e3], device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt)


res = {}
for name, fn_ar in (("graph+nccl", ar), ("graph, no allreduce", None)):
    for _ in range(3):
        eng.train_step(x, tgt, allreduce=fn_ar)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        eng.train_step(x, tgt, allreduce=fn_ar)
    res[name] = timed(gr.replay)
res["eager+nccl"] = timed(lambda: eng.train_step(x, tgt, allreduce=ar), n=100)
# the all-reduce alone, back to back
buf = eng._gradbuf
res["allreduce alone (us)"] = timed(lambda: ar(buf), n=500) * 1e3
if rank == 0:
    print({k: round(v, 4) for k, v in res.items()}, "world", world, "grad floats", buf.numel(), flush=True)
dist.destroy_process_group()
