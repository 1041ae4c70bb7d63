"""`FusedAdam`: the reference's optimizer (`torch.optim.Adam(model.parameters(), lr, weight_decay)`,
/root/reference/Main.py:192; stepped at /root/reference/Helpers/TrainTestHelper.py:142-143) as ONE
multi-tensor kernel of libihgnn_b200.so (`ihg_adam_step`, csrc/adam.cu).

Drop-in for `torch.optim.Adam` on fp32 CUDA parameters with dense gradients: same constructor
arguments (lr, betas, eps, weight_decay), same `param_groups` -- the learning-rate decay rule of
TrainTestHelper.py:155-159 edits `param_group['lr']` and is honoured -- and the same `state_dict()`
layout (`step` fp32 scalar tensor, `exp_avg`, `exp_avg_sq` per parameter), so checkpoints written by
Main.py:252-262 load into either optimizer.  The update is bit-identical with torch's fused CUDA Adam
(tests/test_gpu_parity.py::test_fused_adam_*).  Step counters and the learning rate live on the device,
hence the step is CUDA-graph capturable without a `capturable=` switch.  CUDA only, no fallback."""
from __future__ import annotations

from typing import Iterable

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):

    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("FusedAdam: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._lr_dev = {}            # group index -> (device fp32 scalar, the host value it holds)

    def _init_state(self, p: torch.Tensor) -> dict:
        st = self.state[p]
        if not st:
            _lib.require_cuda(p)
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("FusedAdam: parameters must be contiguous float32 CUDA tensors")
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def refresh_lr(self) -> None:
        """Push `param_group['lr']` to the device scalar the kernel reads, if it changed.  Called by
        `step()`; callers that REPLAY a captured step (graphs.GraphedTrainStep) call it before the replay."""
        for gi, group in enumerate(self.param_groups):
            lr = float(group["lr"])
            held = self._lr_dev.get(gi)
            if held is None:
                dev = next((p.device for p in group["params"]), None)
                if dev is None:
                    continue
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("FusedAdam: run one eager step (or refresh_lr()) before capturing")
                self._lr_dev[gi] = [torch.full((), lr, dtype=torch.float32, device=dev), lr]
            elif held[1] != lr:
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("FusedAdam: the learning rate cannot change inside a capture")
                held[0].fill_(lr)
                held[1] = lr

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if not torch.cuda.is_current_stream_capturing():
            self.refresh_lr()
        for gi, group in enumerate(self.param_groups):
            todo = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad
                if g.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                if not g.is_contiguous() or g.dtype != torch.float32:
                    g = g.contiguous().float()
                todo.append((p, g, self._init_state(p)))
            if not todo:
                continue
            table = (_lib.IhgAdamTensor * len(todo))()
            for k, (p, g, st) in enumerate(todo):
                table[k].param, table[k].grad = _lib.ptr(p), _lib.ptr(g)
                table[k].exp_avg, table[k].exp_avg_sq = _lib.ptr(st["exp_avg"]), _lib.ptr(st["exp_avg_sq"])
                table[k].step, table[k].numel = _lib.ptr(st["step"]), p.numel()
            b1, b2 = group["betas"]
            n_elem = sum(p.numel() for p, _, _ in todo)
            _lib.call("ihg_adam_step", table, len(todo), _lib.ptr(self._lr_dev[gi][0]), float(b1), float(b2),
                      float(group["eps"]), float(group["weight_decay"]), _lib.stream_ptr(),
                      tag="adam_step", algo_bytes=28 * n_elem)
        return loss


def make_adam(params, lr: float = 1e-3, weight_decay: float = 0.0, capturable: bool = False) -> FusedAdam:
    """The reference's `torch.optim.Adam(params, lr=lr, weight_decay=weight_decay)` (Main.py:192) on this
    library's kernel; `capturable` is accepted for signature compatibility (always capturable)."""
    return FusedAdam(params, lr=lr, weight_decay=weight_decay)
