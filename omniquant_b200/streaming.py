"""Host-DRAM block streaming: decoder blocks live in (pinned) host memory and visit the GPU one at a time.

Reference: quantize/omniquant.py:159 (``layer = layers[i].to(dev)``) and :247-250 (``layers[i] = qlayer.to("cpu")``)
move each block synchronously from pageable memory, on the critical path of every block.  Here the NEXT block's
tensors are pinned and copied on a dedicated copy stream while the current block calibrates (its optimiser steps are
already queued on the compute stream), so the H2D transfer costs no wall-clock time; needed for BASELINE config 5
(Falcon-180B: 4.4 GB fp16 per block over PCIe).  Only torch device memory / streams are used -- plumbing, no kernels.
"""
from typing import Optional

import torch
from torch import nn


class BlockStreamer:
    def __init__(self, layers, device, enabled: bool = True):
        self.layers = layers
        self.dev = torch.device(device)
        self.enabled = enabled and self.dev.type == "cuda"
        self.stream = torch.cuda.Stream(device=self.dev) if self.enabled else None
        self._pending = {}            # index -> (module already holding device tensors, event)

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
