"""Whole-forward CUDA-graph capture for latency-bound (small batch) inference — SURVEY.md §8(f)-4.

The reference's timing harness (test_time.py:4-9) runs 10 000 forwards of ONE clip; at batch 1 the ~170 kernel
launches of a forward are launch-bound (the Python + driver cost of a launch exceeds the kernel's run time).
`GraphedForward` captures `model(x)` once into a CUDA graph (torch's capture machinery around the plain CUDA
launches libistvt_b200.so issues on the current stream) and replays it with one `cudaGraphLaunch` per forward.
The packed-weight cache must be warm before the capture (a warm-up forward does that); any parameter update
invalidates the graph — re-create it.
"""
from __future__ import annotations

import torch


class GraphedForward:
    def __init__(self, model, example: torch.Tensor, warmup: int = 3):
        if model.training:
            raise ValueError("CUDA-graph capture is for eval-mode inference")
        if not example.is_cuda:
            raise ValueError("ISTVT (istvt_b200) runs on CUDA tensors only: there is no CPU fallback by design")
        self.model = model
        self.static_in = example.detach().clone().float().contiguous()
        side = torch.cuda.Stream(example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):                       # packs weights, sizes the allocator, sets kernel attributes
                model(self.static_in)
        torch.cuda.current_stream(example.device).wait_stream(side)
        torch.cuda.synchronize(example.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = model(self.static_in)

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """x: same shape as the example (device or pinned host tensor).  Returns the graph's static output tensor
        (overwritten by the next call)."""
        if tuple(x.shape) != tuple(self.static_in.shape):
            raise ValueError(f"graph was captured for {tuple(self.static_in.shape)}, got {tuple(x.shape)}")
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out
