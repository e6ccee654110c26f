"""Host-side clip feeder for inference: pinned host batches -> logits on the host, copies overlapped with compute.

The reference's evaluation loop (train_CNN.py:928-944: `image.cuda()`, `model(image)`, `preds = clas > 0`) pays
the host->device copy of every batch serially before the forward.  `ClipStream` keeps the same per-batch
contract (every batch is copied H2D from pinned memory, its logits are copied D2H) but issues the copy of
batch i+1 on a side stream while batch i computes, with two device staging buffers and CUDA events for the
hand-off.  Nothing here computes: the forward is `model(x)`, i.e. the CUDA path behind the C ABI.
"""
from __future__ import annotations

from typing import Iterable, Iterator

import torch


class ClipStream:
    def __init__(self, model, depth: int = 2):
        self.model = model
        self.depth = depth
        self._bufs = None
        self._key = None

    def _setup(self, shape, device, dtype=torch.float32):
        key = (tuple(shape), str(device), dtype)
        if self._key == key:
            return
        self._key = key
        self._copy = torch.cuda.Stream(device)
        self._bufs = [torch.empty(shape, dtype=dtype, device=device) for _ in range(self.depth)]
        self._ready = [torch.cuda.Event() for _ in range(self.depth)]    # H2D of the slot finished
        self._free = [torch.cuda.Event() for _ in range(self.depth)]     # forward finished reading the slot
        self._done = [torch.cuda.Event() for _ in range(self.depth)]     # logits of the slot are on the host
        self._out = [torch.empty(shape[0], 1, dtype=torch.float32).pin_memory() for _ in range(self.depth)]

    def _prefetch(self, slot: int, host_batch: torch.Tensor, used_before: bool):
        if not host_batch.is_pinned():
            raise ValueError("ClipStream expects pinned host batches (torch.Tensor.pin_memory())")
        with torch.cuda.stream(self._copy):
            if used_before:
                self._copy.wait_event(self._free[slot])
            self._bufs[slot].copy_(host_batch, non_blocking=True)
            self._ready[slot].record(self._copy)

    @torch.no_grad()
    def run(self, host_batches: Iterable[torch.Tensor], device=None) -> Iterator[torch.Tensor]:
        """Yields the logits [B, 1] of every batch, in order (a pinned host tensor that is reused: copy it out
        before asking for the next one)."""
        device = torch.device(device) if device is not None else next(self.model.parameters()).device
        it = iter(host_batches)
        try:
            nxt = next(it)
        except StopIteration:
            return
        self._setup(nxt.shape, device, nxt.dtype)     # fp32 [B,T,3,H,W] or uint8 [B,T,H,W,3] batches
        main = torch.cuda.current_stream(device)
        self._prefetch(0, nxt, used_before=False)
        i = 0
        pending = []
        while nxt is not None:
            slot = i % self.depth
            try:
                after = next(it)
            except StopIteration:
                after = None
            if after is not None:
                if tuple(after.shape) != self._key[0] or after.dtype != self._key[2]:
                    raise ValueError("ClipStream: all batches of one run must have the same shape and dtype")
                self._prefetch((i + 1) % self.depth, after, used_before=(i + 1) >= self.depth)
            main.wait_event(self._ready[slot])
            logits = self.model(self._bufs[slot])
            self._free[slot].record(main)
            self._out[slot].copy_(logits, non_blocking=True)
            self._done[slot].record(main)
            pending.append(slot)
            if len(pending) >= self.depth:            # hand back the oldest result while the newest computes
                s = pending.pop(0)
                self._done[s].synchronize()
                yield self._out[s]
            nxt = after
            i += 1
        for s in pending:
            self._done[s].synchronize()
            yield self._out[s]
