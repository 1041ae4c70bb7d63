"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): the product's partition planner
(ihgnn_b200.dist.PartitionPlan) + the exchange choreography, replayed with CPU ops
(oracle/dist_oracle.py), must reproduce the single-process oracle exactly -- outputs, input
gradients and (all-reduced) weight gradients."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import REPO


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, ret):
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    from ihgnn_b200 import synth
    from ihgnn_b200.dist import PartitionPlan
    from oracle import dist_oracle as dorc
    from oracle import ihgnn_oracle as orc
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        U, Q, I, E, d = 61, 13, 47, 900, 8
        log = synth.make_search_log(U, Q, I, E, 30, shape="cikm", seed=5, zipf=0.9)
        plan = PartitionPlan(log.pos_user, log.pos_query, log.pos_item, U, Q, I, world, rank)
        gen = torch.Generator().manual_seed(3)
        x = torch.randn(U + Q + I, d, generator=gen, dtype=torch.float64)
        layers = []
        for k, order in enumerate([3, 1]):
            K = 7 if order == 3 else 3
            layers.append(dict(order=order,
                               tw=torch.randn(d, d, generator=gen, dtype=torch.float64) / d ** 0.5,
                               tb=torch.randn(d, generator=gen, dtype=torch.float64) * 0.1,
                               aw=torch.randn(d, K * d, generator=gen, dtype=torch.float64) / (K * d) ** 0.5,
                               ab=torch.randn(d, generator=gen, dtype=torch.float64) * 0.1))
        # ---- single-process oracle on the global graph
        g = orc.build_hypergraph(log.pos_user, log.pos_query, log.pos_item, U, Q, I)
        adj = g.adjacency(torch.float64)
        dv = g.VertexDegrees.pow(-1).double()
        xg = x.clone().requires_grad_(True)
        pg = [{k: (v.clone().requires_grad_(True) if torch.is_tensor(v) else v) for k, v in L.items()} for L in layers]
        h = xg
        outs_g = []
        for L in pg:
            h = orc.ihgnn_layer(h, g, adj, dv, L["tw"], L["tb"], L["aw"], L["ab"], L["order"])
            outs_g.append(h)
        torch.cat(outs_g, 1).sum().backward()
        # ---- sharded replay
        own = torch.from_numpy(plan.own_global_ids())
        xo = x[own].clone().requires_grad_(True)
        ps = [{k: (v.clone().requires_grad_(True) if torch.is_tensor(v) else v) for k, v in L.items()} for L in layers]
        h = xo
        outs_s = []
        for L in ps:
            h = dorc.sharded_ihgnn_layer(h, plan, L["tw"], L["tb"], L["aw"], L["ab"], L["order"])
            outs_s.append(h)
        torch.cat(outs_s, 1).sum().backward()
        dorc.allreduce_grads([v for L in ps for v in L.values() if torch.is_tensor(v)])
        err = 0.0
        for a, b in zip(outs_s, outs_g):
            err = max(err, float((a - b[own]).abs().max() / b.abs().max()))
        err = max(err, float((xo.grad - xg.grad[own]).abs().max() / xg.grad.abs().max()))
        for Ls, Lg in zip(ps, pg):
            for k in ("tw", "tb", "aw", "ab"):
                err = max(err, float((Ls[k].grad - Lg[k].grad).abs().max() / Lg[k].grad.abs().max()))
        # plan invariants
        assert plan.edge_count == int((np.searchsorted(plan.ub, log.pos_user, side="right") - 1 == rank).sum())
        assert plan.i3_local.max(initial=-1) < plan.n_local
        tot = torch.tensor([plan.edge_count, plan.n_own], dtype=torch.int64)
        dist.all_reduce(tot)
        assert int(tot[0]) == E and int(tot[1]) == U + Q + I
        ret[rank] = err
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_conv_matches_single_process_oracle(world):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    assert max(ret.values()) < 1e-12, dict(ret)


def test_partition_plan_single_rank_is_identity():
    from ihgnn_b200 import synth
    from ihgnn_b200.dist import PartitionPlan
    log = synth.make_search_log(20, 5, 15, 100, 10, seed=1)
    p = PartitionPlan(log.pos_user, log.pos_query, log.pos_item, 20, 5, 15, 1, 0)
    assert p.n_local == p.n_own == 40 and p.R == 0 and p.S == 0 and p.edge_count == 100
    assert np.array_equal(p.i3_local[:, 0], log.pos_user)
    assert np.array_equal(p.i3_local[:, 1], log.pos_query + 20)
    assert np.array_equal(p.i3_local[:, 2], log.pos_item + 25)


def _fetch_worker(rank: int, world: int, port: int, ret):
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    from ihgnn_b200 import synth
    from ihgnn_b200.dist import PartitionPlan
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        U, Q, I, E, D = 61, 13, 47, 900, 6
        log = synth.make_search_log(U, Q, I, E, 30, shape="cikm", seed=5, zipf=0.9)
        plan = PartitionPlan(log.pos_user, log.pos_query, log.pos_item, U, Q, I, world, rank)
        gen = torch.Generator().manual_seed(7)
        feat = torch.randn(U + Q + I, D, generator=gen, dtype=torch.float64)      # the global matrix
        f_own = feat[torch.from_numpy(plan.own_global_ids())]
        B = 37
        users = torch.randint(0, U, (B,), generator=gen)
        queries = torch.randint(0, Q, (B,), generator=gen)
        items = torch.randint(0, I, (B,), generator=gen)
        local_rows, mine = plan.batch_rows(users, queries, items)
        assert local_rows.shape == mine.shape == (3 * B,) and int(local_rows.min()) >= 0 \
            and int(local_rows.max()) < plan.n_own
        # the choreography of _FetchRowsFn.forward with CPU ops: gather, mask, all-reduce
        rows = f_own[local_rows] * mine.to(torch.float64).view(-1, 1)
        dist.all_reduce(rows)
        want = torch.cat([feat[users], feat[queries + U], feat[items + U + Q]])
        ok_fwd = torch.equal(rows, want)
        # every requested id is owned by exactly one rank
        cnt = mine.to(torch.int64)
        dist.all_reduce(cnt)
        ok_own = bool((cnt == 1).all())
        # backward: masked scatter-add of the (replicated) gradient == the owner's slice of the global one
        gout = torch.randn(3 * B, D, generator=gen, dtype=torch.float64)
        df = torch.zeros_like(f_own).index_add_(0, local_rows, gout * mine.to(torch.float64).view(-1, 1))
        gidx = torch.cat([users, queries + U, items + U + Q])
        dglobal = torch.zeros_like(feat).index_add_(0, gidx, gout)
        ok_bwd = torch.allclose(df, dglobal[torch.from_numpy(plan.own_global_ids())], rtol=0, atol=1e-12)
        # a second call with the same batch size reuses the cached range tensors
        lr2, m2 = plan.batch_rows(users, queries, items)
        ret[rank] = (bool(ok_fwd), ok_own, bool(ok_bwd), bool(torch.equal(lr2, local_rows) and torch.equal(m2, mine)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_fixed_shape_batch_row_fetch(world):
    """PartitionPlan.batch_rows + the gather / mask / all-reduce choreography of the sharded batch head
    (dist._FetchRowsFn) reproduce the global row selects of RawGnn.py:128-133 and their backward."""
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_fetch_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world and all(all(v) for v in ret.values()), dict(ret)
