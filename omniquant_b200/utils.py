"""Step / logging helpers with the reference's names (reference utils.py:11-79).

The B200 path stores activations in bf16, whose exponent range equals fp32's, so no loss scaling is
needed: ``NativeScalerWithGradNormCount`` keeps the reference's call signature (scaled backward,
gradient norm, optimiser step) with a scale fixed at 1."""
import logging
import os
import sys
import time
from math import inf

import torch


def ampscaler_get_grad_norm(parameters, norm_type: float = 2.0) -> torch.Tensor:
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad.detach() for p in parameters if p.grad is not None]
    if not grads:
        return torch.tensor(0.0)
    dev = grads[0].device
    if float(norm_type) == inf:
        return max(g.abs().max().to(dev) for g in grads)
    return torch.norm(torch.stack([torch.norm(g, float(norm_type)).to(dev) for g in grads]), float(norm_type))


class NativeScalerWithGradNormCount:
    state_dict_key = "amp_scaler"

    def __init__(self):
        self._scale = 1.0

    def __call__(self, loss, optimizer, clip_grad=None, parameters=None, create_graph=False, update_grad=True,
                 retain_graph=False):
        loss.backward(create_graph=create_graph, retain_graph=retain_graph)
        if not update_grad:
            return None
        if clip_grad is not None:
            assert parameters is not None
            norm = torch.nn.utils.clip_grad_norm_(parameters, clip_grad)
        else:
            norm = ampscaler_get_grad_norm(parameters)
        optimizer.step()
        return norm

    def state_dict(self):
        return {"scale": self._scale}

    def load_state_dict(self, state_dict):
        self._scale = state_dict.get("scale", 1.0)


def create_logger(output_dir, dist_rank=0, name=""):
    logger = logging.getLogger(name)
    logger.setLevel(logging.INFO)
    logger.propagate = False
    fmt = "[%(asctime)s %(name)s] (%(filename)s %(lineno)d): %(levelname)s %(message)s"
    if dist_rank == 0 and not any(isinstance(h, logging.StreamHandler) for h in logger.handlers):
        console = logging.StreamHandler(sys.stdout)
        console.setLevel(logging.DEBUG)
        console.setFormatter(logging.Formatter(fmt=fmt, datefmt="%Y-%m-%d %H:%M:%S"))
        logger.addHandler(console)
    if output_dir:
        fh = logging.FileHandler(os.path.join(output_dir, f"log_rank{dist_rank}_{int(time.time())}.txt"), mode="a")
        fh.setLevel(logging.DEBUG)
        fh.setFormatter(logging.Formatter(fmt=fmt, datefmt="%Y-%m-%d %H:%M:%S"))
        logger.addHandler(fh)
    return logger
