"""Sample-sharded data parallelism of the calibration loop.

The path shards by calibration sample: block weights and LWC/LET parameters are replicated, every rank runs the
same fold + weight fake-quant, forward/backward on its own sample(s), then ONE all-reduce (sum) of the flat
LWC+LET gradient buffer; the fused AdamW scales by 1/world so every rank takes the identical step.  This equals
the reference run with ``--batch_size world`` (mean-reduced MSE over the batch, quantize/omniquant.py:142,214-222).
One process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from typing import Callable, List, Optional

import torch


def shard_indices(nsamples: int, global_batch: int, rank: int, world: int) -> List[int]:
    """Samples owned by ``rank``: global step j consumes samples [j*B, (j+1)*B); the rank takes its b = B/world
    consecutive ones of every step, in step order."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not a multiple of the world size {world}")
    if nsamples % global_batch:
        raise ValueError(f"nsamples {nsamples} is not a multiple of the global batch {global_batch}")
    b = global_batch // world
    return [j * global_batch + rank * b + i for j in range(nsamples // global_batch) for i in range(b)]


def make_allreduce(group=None) -> Optional[Callable[[torch.Tensor], None]]:
    """Sum-all-reduce closure over the default (or given) process group, or None for a single process."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    return lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def world_info():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1
