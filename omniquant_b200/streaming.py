"""Host-DRAM block streaming: decoder blocks live in (pinned) host memory and visit the GPU one at a time.

Reference: quantize/omniquant.py:159 (``layer = layers[i].to(dev)``) and :247-250 (``layers[i] = qlayer.to("cpu")``)
move each block synchronously from pageable memory, on the critical path of every block.  Here the NEXT block's
tensors are pinned and copied on a dedicated copy stream while the current block calibrates (its optimiser steps are
already queued on the compute stream), so the H2D transfer costs no wall-clock time; needed for BASELINE config 5
(Falcon-180B: 4.4 GB fp16 per block over PCIe).  The calibrated block goes back the same way: ``writeback`` copies its
tensors into pinned staging on the copy stream and a host thread moves them to their final (pageable) home while the next
block calibrates.  Only torch device memory / streams are used -- plumbing, no kernels.
"""
import threading
from typing import Optional

import torch
from torch import nn

_STAGING = []       # pinned staging arenas (uint8), re-used across blocks and across omniquant() calls
_STAGING_LOCK = threading.Lock()


def _take_arena(nbytes: int) -> torch.Tensor:
    with _STAGING_LOCK:
        for i, a in enumerate(_STAGING):
            if a.numel() >= nbytes:
                return _STAGING.pop(i)
        _STAGING.clear()       # block shapes changed: drop the smaller arenas
    return torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)


def _give_arena(a: torch.Tensor) -> None:
    with _STAGING_LOCK:
        if len(_STAGING) < 2:
            _STAGING.append(a)


class BlockStreamer:
    def __init__(self, layers, device, enabled: bool = True):
        self.layers = layers
        self.dev = torch.device(device)
        self.enabled = enabled and self.dev.type == "cuda"
        self.stream = torch.cuda.Stream(device=self.dev) if self.enabled else None
        self._pending = {}            # index -> (module already holding device tensors, event)
        self._writers = []            # host threads finishing asynchronous write-backs

    def prefetch(self, i: int) -> None:
        """Start the asynchronous host->device copy of block i (no-op if out of range / disabled / already started)."""
        if not self.enabled or i >= len(self.layers) or i in self._pending:
            return
        layer = self.layers[i]
        for t in list(layer.parameters()) + list(layer.buffers()):
            if t.device.type == "cpu" and not t.is_pinned():
                t.data = t.data.pin_memory()
        with torch.cuda.stream(self.stream):
            layer = layer.to(self.dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._pending[i] = (layer, ev)

    def fetch(self, i: int) -> nn.Module:
        """Block i on the device; waits (on the compute stream, not the host) for a pending prefetch."""
        if i in self._pending:
            layer, ev = self._pending.pop(i)
            cur = torch.cuda.current_stream(self.dev)
            cur.wait_event(ev)
            for t in list(layer.parameters()) + list(layer.buffers()):
                t.data.record_stream(cur)      # allocated on the copy stream, consumed on the compute stream
            return layer
        return self.layers[i].to(self.dev)

    def bytes_of(self, i: int) -> int:
        layer = self.layers[i]
        return sum(t.numel() * t.element_size() for t in list(layer.parameters()) + list(layer.buffers()))

    def writeback(self, module: nn.Module) -> nn.Module:
        """``module.to("cpu")`` (quantize/omniquant.py:248) without stalling the GPU: device tensors are copied into pinned
        staging on the copy stream (after everything queued on the compute stream so far), the module's parameters and
        buffers are re-pointed at their final host tensors at once, and a host thread fills those in when the copy has
        landed.  ``drain()`` must be called before the host tensors are read."""
        if not self.enabled:
            return module.to("cpu")
        tensors = [t for t in list(module.parameters()) + list(module.buffers()) if t.device.type == "cuda"]
        if not tensors:
            return module
        align = lambda n: (n + 255) // 256 * 256
        total = sum(align(t.numel() * t.element_size()) for t in tensors)
        arena = _take_arena(total)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.dev))
        jobs, off = [], 0
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            for t in tensors:
                src = t.data
                nb = src.numel() * src.element_size()
                stage = arena[off:off + nb].view(src.dtype).view(src.shape)
                off += align(nb)
                stage.copy_(src.contiguous(), non_blocking=True)
                src.record_stream(self.stream)
                home = torch.empty(src.shape, dtype=src.dtype)
                jobs.append((stage, home))
                t.data = home
            done = torch.cuda.Event()
            done.record(self.stream)

        def finish():
            done.synchronize()
            for stage, home in jobs:
                home.copy_(stage)
            _give_arena(arena)

        th = threading.Thread(target=finish, daemon=True)
        th.start()
        self._writers.append(th)
        return module

    def drain(self) -> None:
        """Wait for every pending write-back (call before the host copies of the calibrated blocks are used)."""
        for th in self._writers:
            th.join()
        self._writers = []
