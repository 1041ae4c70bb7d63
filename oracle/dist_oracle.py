"""CPU (gloo) restatement of the multi-rank hypergraph-convolution choreography.

TEST INFRASTRUCTURE ONLY.  It consumes the product's host-side planner
(ihgnn_b200.dist.PartitionPlan) and replays the exchange steps of ihgnn_b200/dist.py
(halo_exchange -> local node->edge->node math -> halo_reduce) with plain torch CPU ops and
torch.distributed point-to-point messages, so that world_size>1 runs can be checked on CPU
against the single-process oracle (oracle/ihgnn_oracle.py).  The local math is the oracle's
own (`feature_interactor`, index_add); nothing here is used by the product.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist

from . import ihgnn_oracle as orc


def _a2a_rows(inp: torch.Tensor, in_counts: List[int], out_counts: List[int]) -> torch.Tensor:
    """all-to-all of row blocks via isend/irecv (gloo has no all_to_all_single)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    d = inp.shape[1]
    outs = [torch.empty((c, d), dtype=inp.dtype) for c in out_counts]
    ins = list(torch.split(inp.contiguous(), in_counts))
    ops = []
    for peer in range(world):
        if peer == rank:
            outs[peer].copy_(ins[peer])
            continue
        if in_counts[peer]:
            ops.append(dist.P2POp(dist.isend, ins[peer].contiguous(), peer))
        if out_counts[peer]:
            ops.append(dist.P2POp(dist.irecv, outs[peer], peer))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return torch.cat(outs) if outs else inp.new_zeros((0, d))


class _HaloExchange(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_own, plan):
        ctx.plan = plan
        send = x_own[torch.from_numpy(plan.send_rows)]
        recv = _a2a_rows(send, [int(c) for c in plan.send_counts], [int(c) for c in plan.recv_counts])
        return torch.cat([x_own, recv])              # local layout == [own rows ; receive layout]

    @staticmethod
    def backward(ctx, dx_local):
        return _halo_reduce(dx_local, ctx.plan), None


def _halo_reduce(s_local: torch.Tensor, plan) -> torch.Tensor:
    """own rows <- own partial + partials received from the ranks that hold them as halo rows,
    summed in the plan's fixed order (own first, then ascending source rank)."""
    send = s_local[plan.n_own:]
    recv = _a2a_rows(send, [int(c) for c in plan.recv_counts], [int(c) for c in plan.send_counts])
    out = s_local[:plan.n_own].clone()
    rowptr, col = plan.reduce_rowptr, plan.reduce_col
    rows = torch.repeat_interleave(torch.arange(plan.n_own), torch.from_numpy(rowptr[1:] - rowptr[:-1]))
    out.index_add_(0, rows, recv[torch.from_numpy(col)])
    return out


class _HaloReduce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s_local, plan):
        ctx.plan = plan
        return _halo_reduce(s_local, plan)

    @staticmethod
    def backward(ctx, dout_own):
        return _HaloExchange.apply(dout_own, ctx.plan), None


def sharded_ihgnn_layer(x_own: torch.Tensor, plan, transform_weight, transform_bias, agg_weight,
                        agg_bias, order: int) -> torch.Tensor:
    """One IHGNN layer (Models/GnnLayers.py:221-236) on the own rows of this rank."""
    xp_own = torch.nn.functional.linear(x_own, transform_weight, transform_bias)
    xp = _HaloExchange.apply(xp_own, plan)                                  # [n_local, d]
    i3 = torch.from_numpy(plan.i3_local)
    ef = orc.feature_interactor(xp, i3, agg_weight, agg_bias, order)        # local hyperedges
    s_local = torch.zeros((plan.n_local, x_own.shape[1]), dtype=x_own.dtype)
    for s in range(3):
        s_local = s_local.index_add(0, i3[:, s], ef)
    s_own = _HaloReduce.apply(s_local, plan)
    return torch.from_numpy(plan.dv_inv_own).to(x_own.dtype).view(-1, 1) * s_own


def allreduce_grads(params) -> None:
    for p in params:
        if p.grad is not None:
            dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
