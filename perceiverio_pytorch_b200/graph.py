"""CUDA-graph capture of a hot-path forward.

Every C-ABI entry point only enqueues work on the caller's stream (no host synchronisation, TMA descriptors are
passed by value as kernel parameters), so a whole PerceiverEncoder + PerceiverDecoder forward — several hundred
launches for the 48-layer ImageNet recipe — can be captured once and replayed without any host-side launch cost
(SURVEY.md §7-H6).
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class GraphedForward:
    """Capture `fn(*inputs)` (tensors in, tensor or tuple of tensors out) for fixed shapes.

    `__call__(*inputs)` copies the arguments into the captured input buffers (skipped for arguments that already
    are those buffers — see `.inputs`), replays the graph and returns the captured output tensors (overwritten by
    the next call)."""

    def __init__(self, fn: Callable, example_inputs: Sequence[torch.Tensor], warmup: int = 2):
        self.inputs = [t.clone() if isinstance(t, torch.Tensor) else t for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.inference_mode():
            for _ in range(max(1, warmup)):  # builds the derived-weight caches and sets kernel attributes
                fn(*self.inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.inference_mode(), torch.cuda.graph(self.graph):
            self.outputs = fn(*self.inputs)

    def __call__(self, *inputs):
        for dst, src in zip(self.inputs, inputs):
            if isinstance(dst, torch.Tensor) and src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.outputs
