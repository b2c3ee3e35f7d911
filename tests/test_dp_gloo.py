"""world_size-2 data parallelism on CPU (gloo): the host-side sharding logic and the DP algebra the GPU path
relies on -- per-rank gradients on the rank's own sample, summed by one all-reduce of the flat buffer and
scaled by 1/world, equal the single-process gradient of the reference run with --batch_size 2; AdamW on the
reduced buffer keeps the replicas bit-identical.  (The oracle computes the per-rank gradients here; the CUDA
engine is exercised at N>1 by bench.py --gpus N on the GPU box.)"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from golden_util import BlockFixture
from omniquant_b200.parallel import shard_indices
from oracle import oq_oracle as orc


def test_shard_indices_partition():
    for n, B, w in ((128, 8, 8), (128, 2, 2), (16, 4, 2), (8, 1, 1)):
        seen = []
        for r in range(w):
            idx = shard_indices(n, B, r, w)
            assert len(idx) == n // w
            seen += idx
        assert sorted(seen) == list(range(n))
        # step j of every rank together = the reference's batch j
        b = B // w
        for j in range(n // B):
            step = sorted(i for r in range(w) for i in shard_indices(n, B, r, w)[j * b:(j + 1) * b])
            assert step == list(range(j * B, (j + 1) * B))
    with pytest.raises(ValueError):
        shard_indices(128, 3, 0, 2)


def _worker(rank, world, port, fname, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    fx = BlockFixture(fname)
    cfg = orc.QCfg(fx.wbits, fx.abits, fx.group_size, False, fx.lwc, fx.let)
    Wt, P = fx.weights(), fx.init_params()
    keys = sorted(P)
    for k in keys:
        P[k].requires_grad_(True)
    flat = torch.zeros(sum(P[k].numel() for k in keys))
    idx = shard_indices(fx.nsamples, world, rank, world)
    cos, sin = orc.rope_tables(fx.H // fx.heads, fx.T)
    mask = orc.causal_mask(fx.T)
    fwd = lambda x, T, aq: orc.llama_forward(x, T, fx.heads, cos, sin, mask, cfg, aq)
    with torch.no_grad():
        Tfp = orc.llama_prepare(Wt, P, orc.QCfg(fx.wbits, fx.abits, fx.group_size, False, fx.lwc, False), quant=False)
    opt = torch.optim.AdamW([P[k] for k in keys], lr=1e-2, weight_decay=0)
    losses = []
    for j in idx[:2]:                                   # two optimiser steps
        x = fx.inps[j:j + 1]
        with torch.no_grad():
            tgt = fwd(x, Tfp, False)
        loss = torch.nn.functional.mse_loss(tgt, fwd(x, orc.llama_prepare(Wt, P, cfg), True))
        opt.zero_grad()
        loss.backward()
        o = 0
        for k in keys:                                  # one flat buffer, one collective, then 1/world
            flat[o:o + P[k].numel()] = P[k].grad.reshape(-1)
            o += P[k].numel()
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        lt = loss.detach().clone()
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        flat /= world
        o = 0
        for k in keys:
            P[k].grad.copy_(flat[o:o + P[k].numel()].view_as(P[k]))
            o += P[k].numel()
        opt.step()
        losses.append(float(lt) / world)
    torch.save({"P": {k: P[k].detach() for k in keys}, "losses": losses}, f"{out}.{rank}")
    dist.destroy_process_group()


def test_two_rank_dp_equals_reference_batch2(tmp_path):
    fname = "llama_w4a4_lwc_let_bs2.npz"
    out = str(tmp_path / "res")
    port = 29000 + os.getpid() % 2000
    mp.start_processes(_worker, args=(2, port, fname, out), nprocs=2, join=True, start_method="spawn")
    r0, r1 = torch.load(f"{out}.0"), torch.load(f"{out}.1")
    for k in r0["P"]:
        assert torch.equal(r0["P"][k], r1["P"][k]), f"replicas diverged on {k}"
    # the reference itself with --batch_size 2 (fixture): same first-step loss (later steps use its two learning rates)
    fx = BlockFixture(fname)
    np.testing.assert_allclose(r0["losses"][0], fx.losses[0], rtol=2e-4)
    # and the same parameters as one process stepping on the batch of 2 with the same (single-lr) optimiser
    cfg = orc.QCfg(fx.wbits, fx.abits, fx.group_size, False, fx.lwc, fx.let)
    Wt, P = fx.weights(), fx.init_params()
    keys = sorted(P)
    for k in keys:
        P[k].requires_grad_(True)
    cos, sin = orc.rope_tables(fx.H // fx.heads, fx.T)
    mask = orc.causal_mask(fx.T)
    fwd = lambda x, T, aq: orc.llama_forward(x, T, fx.heads, cos, sin, mask, cfg, aq)
    with torch.no_grad():
        Tfp = orc.llama_prepare(Wt, P, orc.QCfg(fx.wbits, fx.abits, fx.group_size, False, fx.lwc, False), quant=False)
    opt = torch.optim.AdamW([P[k] for k in keys], lr=1e-2, weight_decay=0)
    for j in range(2):
        x = fx.inps[2 * j:2 * j + 2]
        with torch.no_grad():
            tgt = fwd(x, Tfp, False)
        loss = torch.nn.functional.mse_loss(tgt, fwd(x, orc.llama_prepare(Wt, P, cfg), True))
        opt.zero_grad()
        loss.backward()
        opt.step()
    for k in keys:
        np.testing.assert_allclose(r0["P"][k].numpy(), P[k].detach().numpy(), rtol=0, atol=2e-5, err_msg=k)
