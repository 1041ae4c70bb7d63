"""CUDA-graph capture of the launch-bound steps.

A training step of the reference (`Helpers/TrainTestHelper.py:123-143`: forward, BCE-with-logits,
`loss.backward()`, `optimizer.step()`, `optimizer.zero_grad()`) is ~90 kernel launches of which most
run for a few microseconds (batch gathers, scoring, the per-table gradient scatters, Adam).  Every
entry point of libihgnn_b200.so is asynchronous on the current stream, allocates nothing and never
synchronises (include/ihgnn_b200.h), so the whole step can be captured once and replayed as ONE
graph launch: the host cost per step becomes four small index copies plus the loss read-back.

`GraphedTrainStep` owns static device buffers for the batch; `__call__` copies a (host or device)
batch into them, replays the graph and returns the (static) loss tensor.  The batch shape is fixed at
construction (the reference's batches are fixed-size except the last one of an epoch -- run that one
eagerly).  `graph_callable` captures any argument-less closure (bench.py uses it for the conv-only
metric).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


def graph_callable(fn: Callable[[], None], warmup: int = 3) -> torch.cuda.CUDAGraph:
    """Warm `fn` up on a side stream (lazy plans, cuBLAS handles, allocator pools), then capture one
    call into a CUDA graph.  `fn` must be sync-free and must use static tensors for its inputs."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(max(warmup, 1)):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


class GraphedTrainStep:
    """forward -> BCEWithLogits(mean) -> backward -> optimizer.step(), captured once.

    model(users, queries, items) -> scores [B];  optimizer must be capturable
    (`torch.optim.Adam(params, lr, fused=True, capturable=True)` is the reference's Adam,
    Main.py:192, in capturable form)."""

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, batch_rows: int,
                 device, loss_fn: Optional[Callable] = None, warmup: int = 3,
                 example=None, after_backward: Optional[Callable[[], None]] = None):
        dev = torch.device(device)
        self.model, self.optimizer = model, optimizer
        self.users = torch.zeros(batch_rows, dtype=torch.int64, device=dev)
        self.queries = torch.zeros(batch_rows, dtype=torch.int64, device=dev)
        self.items = torch.zeros(batch_rows, dtype=torch.int64, device=dev)
        self.flags = torch.zeros(batch_rows, dtype=torch.float32, device=dev)
        if example is not None:
            self._load(*example)
        self.loss_fn = loss_fn or torch.nn.functional.binary_cross_entropy_with_logits
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self._after_backward = after_backward

        def step():
            scores = self.model(self.users, self.queries, self.items)
            loss = self.loss_fn(scores, self.flags)
            self.optimizer.zero_grad(set_to_none=True)
            loss.backward()
            if self._after_backward is not None:
                self._after_backward()
            self.optimizer.step()
            self.loss.copy_(loss.detach())

        self.graph = graph_callable(step, warmup)

    def _load(self, users, queries, items, flags) -> None:
        self.users.copy_(users, non_blocking=True)
        self.queries.copy_(queries, non_blocking=True)
        self.items.copy_(items, non_blocking=True)
        self.flags.copy_(flags, non_blocking=True)

    def __call__(self, users, queries, items, flags) -> torch.Tensor:
        self._load(users, queries, items, flags)
        self.graph.replay()
        return self.loss
