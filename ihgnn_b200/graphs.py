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
metric).  Construction has no side effect on the model or the optimizer: the warm-up steps that
precede the capture are undone (parameters and optimizer state are restored in place).
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


def _snapshot(model: torch.nn.Module, optimizer: torch.optim.Optimizer):
    params = [p.detach().clone() for p in model.parameters()]
    had_state = {id(p) for group in optimizer.param_groups for p in group["params"] if p in optimizer.state}
    state = {id(p): {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in optimizer.state[p].items()}
             for group in optimizer.param_groups for p in group["params"] if p in optimizer.state}
    return params, had_state, state


@torch.no_grad()
def _restore(model: torch.nn.Module, optimizer: torch.optim.Optimizer, snap) -> None:
    """Copy the snapshot back IN PLACE (the captured graph holds the addresses of the parameters and
    of the optimizer's state tensors); state the warm-up created is reset to its initial value."""
    params, had_state, state = snap
    for p, v in zip(model.parameters(), params):
        p.copy_(v)
    for group in optimizer.param_groups:
        for p in group["params"]:
            st = optimizer.state.get(p)
            if st is None:
                continue
            old = state.get(id(p))
            for k, v in st.items():
                if not torch.is_tensor(v):
                    if old is not None:
                        st[k] = old[k]
                    elif k == "step":
                        st[k] = 0
                elif old is not None:
                    v.copy_(old[k])
                else:
                    v.zero_()           # exp_avg / exp_avg_sq / step of a fresh Adam
    torch.cuda.synchronize()


class GraphedTrainStep:
    """forward -> BCEWithLogits(mean) -> backward -> optimizer.step(), captured once.

    model(users, queries, items) -> scores [B];  optimizer must be capturable:
    `ihgnn_b200.optim.FusedAdam(params, lr)` (the reference's Adam, Main.py:192, as one kernel of this
    library) or `torch.optim.Adam(params, lr, fused=True, capturable=True)`."""

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
        self.loss_sum = torch.zeros((), dtype=torch.float64, device=dev)      # accumulated on the device, see pop_loss_sum
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
            self.loss_sum.add_(loss.detach())

        # warm-up and capture run real optimizer steps (capture itself executes nothing, the warm-up does):
        # snapshot the parameters and the optimizer state first and put them back afterwards, so that
        # building a GraphedTrainStep leaves the caller's model and optimizer exactly as they were
        snap = _snapshot(model, optimizer)
        self.graph = graph_callable(step, warmup)
        _restore(model, optimizer, snap)
        self.loss_sum.zero_()

    def _load(self, users, queries, items, flags) -> None:
        self.users.copy_(users, non_blocking=True)
        self.queries.copy_(queries, non_blocking=True)
        self.items.copy_(items, non_blocking=True)
        self.flags.copy_(flags, non_blocking=True)

    def __call__(self, users, queries, items, flags) -> torch.Tensor:
        self._load(users, queries, items, flags)
        refresh = getattr(self.optimizer, "refresh_lr", None)
        if refresh is not None:
            refresh()                  # param_group['lr'] edits (TrainTestHelper.py:155-159) reach the device scalar
        self.graph.replay()
        return self.loss

    def pop_loss_sum(self) -> float:
        """Sum of the losses of all steps since the last call (ONE device -> host read): the epoch average of
        TrainTestHelper.py:133-147 without the reference's `loss.item()` synchronisation in every step."""
        total = float(self.loss_sum)
        self.loss_sum.zero_()
        return total
