"""Sample-sharded data parallelism of the CUDA BlockEngine: two engine ranks (two processes sharing the one GPU of the
test box, gloo as the transport of the flat-gradient all-reduce -- NCCL needs one GPU per rank and is exercised by
bench.py --gpus N, which carries the same check as `dp_check`) must
  * end every optimiser step with bit-identical parameters on both ranks (same reduced gradient, same fused AdamW), and
  * follow the single-rank engine run with batch = 2 on the same samples (the reference's --batch_size 2,
    quantize/omniquant.py:214-229): same step-0 loss, same learned parameters up to reduction-order round-off.
"""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

H, F_, HEADS, T, N = 256, 512, 2, 128, 8


def _block(seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).half().float()
    W = {"q": r(H, H, sc=0.03), "k": r(H, H, sc=0.03), "v": r(H, H, sc=0.03), "o": r(H, H, sc=0.03),
         "gate": r(F_, H, sc=0.03), "up": r(F_, H, sc=0.03), "down": r(H, F_, sc=0.03),
         "ln1": 1 + r(H, sc=0.1), "ln2": 1 + r(H, sc=0.1)}
    x = r(N, T, H)
    x[:, :, 3] *= 8.0
    act = {f"{k}_s": r(H).abs() * 2 + 0.1 for k in ("qkv", "out", "fc1")}
    return W, x.to(torch.bfloat16), act


def _engine(W, act, batch, world):
    from omniquant_b200.engine import BlockEngine, QuantConfig
    q = QuantConfig(wbits=4, abits=4, group_size=0, lwc=True, let=True, recip_div=True)
    return BlockEngine("llama", W, n_heads=HEADS, seq_len=T, qcfg=q, eps=1e-6, batch=batch, let_init=act, world_size=world)


def _rank(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from omniquant_b200.engine import fp_pass, train
    from omniquant_b200.parallel import make_allreduce, shard_indices
    torch.cuda.set_device(0)
    W, x, act = _block()
    idx = shard_indices(N, world, rank, world)
    xs = x[idx].cuda()
    eng = _engine(W, act, 1, world)
    fp, qi = xs.clone(), xs.clone()
    fp_pass(eng, fp)
    losses, _ = train(eng, fp, qi, epochs=2, allreduce=make_allreduce(), use_graph=False)
    torch.cuda.synchronize()
    torch.save({"params": eng.params.cpu(), "losses": losses.cpu()}, f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_two_engine_ranks_are_identical_and_equal_batch2(tmp_path):
    from omniquant_b200.engine import fp_pass, train
    out = str(tmp_path / "dp")
    port = 29500 + os.getpid() % 1500
    mp.start_processes(_rank, args=(2, port, out), nprocs=2, join=True, start_method="spawn")
    r0, r1 = torch.load(f"{out}.0"), torch.load(f"{out}.1")
    assert torch.equal(r0["params"], r1["params"]), "replicas diverged"
    assert torch.equal(r0["losses"], r1["losses"])
    # single rank, batch 2, same global sample order
    W, x, act = _block()
    eng = _engine(W, act, 2, 1)
    xs = x.cuda()
    fp, qi = xs.clone(), xs.clone()
    fp_pass(eng, fp)
    losses, _ = train(eng, fp, qi, epochs=2, use_graph=False)
    torch.cuda.synchronize()
    l_dp, l_b2 = r0["losses"], losses.cpu()
    print("dp losses", l_dp, "\nbatch-2 losses", l_b2)
    # step 0 sees identical parameters: the two half-batch losses average to the batch loss
    assert abs(float(l_dp[0]) - float(l_b2[0])) / float(l_b2[0]) < 1e-5
    torch.testing.assert_close(l_dp, l_b2, rtol=2e-2, atol=0)
    # AdamW turns last-bit differences of near-zero gradients into +-lr steps: compare on average
    assert float((r0["params"] - eng.params.cpu()).abs().mean()) < 1e-3
