"""Where one calibrate_block() of bench.py spends its time per rank (run plain for N=1 or under torchrun)."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import bench  # noqa: E402
from omniquant_b200 import engine as E  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
allreduce = (lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM)) if world > 1 else None
marks = []


def mark(name):
    torch.cuda.synchronize()
    marks.append((name, time.perf_counter()))


for fn in ("fp_pass", "train", "requant_pass"):
    orig = getattr(E, fn)

    def wrap(*a, _o=orig, _n=fn, **k):
        mark(f"before {_n}")
        r = _o(*a, **k)
        mark(f"after {_n}")
        return r
    setattr(E, fn, wrap)
_enter, _exit = torch.cuda.graph.__enter__, torch.cuda.graph.__exit__
torch.cuda.graph.__enter__ = lambda self: (mark("capture begin"), _enter(self))[1]
torch.cuda.graph.__exit__ = lambda self, *a: (_exit(self, *a), mark("capture end"))[0]
_replay = torch.cuda.CUDAGraph.replay
state = {"n": 0}


def replay(self):
    state["n"] += 1
    if state["n"] in (1, 2, 3, 10):
        mark(f"before replay {state['n']}")
    return _replay(self)


torch.cuda.CUDAGraph.replay = replay
W_host = bench.block_weights()
W = {k: v.to(dev) for k, v in W_host.items()}
x_dev = bench.calib_inputs(range(rank, bench.NSAMPLES, world)).to(dev)
q = E.QuantConfig(wbits=4, abits=4, group_size=0, lwc=True, let=True, recip_div=True)
init = bench.let_init(W_host)
for it in range(2):
    marks.clear(); state["n"] = 0
    mark("start")
    eng = E.BlockEngine("llama", W, n_heads=bench.HEADS, seq_len=bench.T, qcfg=q, eps=1e-6, device=dev, batch=1, let_init=init,
                        world_size=world)
    mark("engine built")
    fp, qi = x_dev.clone(), x_dev.clone()
    E.calibrate_block(eng, fp, qi, epochs=bench.EPOCHS, allreduce=allreduce, use_graph=True)
    mark("end")
if rank == 0:
    t0 = prev = marks[0][1]
    for n, t in marks:
        print(f"{n:28s} +{(t - prev) * 1e3:9.1f} ms   @{(t - t0) * 1e3:9.1f}", flush=True)
        prev = t
if world > 1:
    dist.destroy_process_group()
