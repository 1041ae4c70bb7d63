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


@pytest.mark.parametrize("world", [3, 5])
def test_routed_reduce_layout_and_work_order(world):
    """Host side of the routed reduce (ihgnn_b200.dist, `ihg_*_routed`): a holder h stores its partial sums of
    the rows owner p owns -- one contiguous chunk of h's local table -- at row send_of[p][:h].sum() of p's
    receive buffer; p's ordered sum (reduce_rowptr / reduce_col over the flat [S rows, by source rank] layout)
    must then see exactly the rows it expects.  Replayed in numpy over the plans of all ranks against the
    global sum; plus the interleaved processing order of the halo work items."""
    from ihgnn_b200 import synth
    from ihgnn_b200.dist import PartitionPlan, halo_work_order
    U, Q, I, E, d = 83, 17, 59, 1500, 4
    log = synth.make_search_log(U, Q, I, E, 30, shape="cikm", seed=8, zipf=0.9)
    plans = [PartitionPlan(log.pos_user, log.pos_query, log.pos_item, U, Q, I, world, r) for r in range(world)]
    N = U + Q + I
    send_of = np.stack([p.send_counts for p in plans])                   # send_of[s][d]
    rng = np.random.default_rng(0)
    # every rank's partial sums over its local rows; global ids of the local rows
    glob = []
    for h, p in enumerate(plans):
        own = p.own_global_ids()
        ids = np.full(p.n_local, -1, dtype=np.int64)
        ids[:p.n_own] = own
        i3g = np.stack([log.pos_user, U + log.pos_query, U + Q + log.pos_item], 1)[p.edge_ids]
        ids[p.i3_local.reshape(-1)] = i3g.reshape(-1)                    # halo rows get their global id from the edges
        assert (ids >= 0).all()
        glob.append(ids)
    part = [rng.standard_normal((p.n_local, d)) for p in plans]
    want = np.zeros((N, d))
    for ids, t in zip(glob, part):
        np.add.at(want, ids, t)
    for pr, p in enumerate(plans):
        recv = np.full((p.S, d), np.nan)
        recv_ids = np.full(p.S, -1, dtype=np.int64)
        for h, q in enumerate(plans):
            if h == pr:
                continue
            c0 = q.n_own + int(q.recv_counts[:pr].sum())                 # h's chunk of rows owned by pr
            n = int(q.recv_counts[pr])
            assert n == send_of[pr][h]
            at = int(send_of[pr][:h].sum())                              # where h's block starts in pr's buffer
            recv[at:at + n] = part[h][c0:c0 + n]
            recv_ids[at:at + n] = glob[h][c0:c0 + n]
        assert not np.isnan(recv).any()
        own_ids = p.own_global_ids()
        assert np.array_equal(recv_ids, own_ids[p.send_rows])           # the flat receive layout IS the send order
        out = part[pr][:p.n_own].copy()
        rp, rc = p.reduce_rowptr, p.reduce_col
        for v in range(p.n_own):
            for j in rc[rp[v]:rp[v + 1]]:
                out[v] += recv[j]
        assert np.allclose(out, want[own_ids], rtol=0, atol=1e-12)
        # work order: a permutation, own rows first and untouched, halo items cycling over the owners
        rows = torch.from_numpy(np.sort(np.concatenate([np.arange(p.n_local), rng.integers(0, p.n_local, 40)])))
        order = halo_work_order(rows, p.n_own, p.recv_counts, world, pr).numpy()
        assert np.array_equal(np.sort(order), np.arange(rows.numel()))
        r = rows.numpy()[order]
        n_own_items = int((rows.numpy() < p.n_own).sum())
        assert np.array_equal(order[:n_own_items], np.arange(n_own_items))
        ends = np.cumsum(p.recv_counts) + p.n_own
        owners = np.searchsorted(ends, r[n_own_items:], side="right")
        k = int(min(np.bincount(owners, minlength=world)[[o for o in range(world) if o != pr and p.recv_counts[o] > 0]]))
        live = [o for o in range(world) if o != pr and p.recv_counts[o] > 0]
        first = owners[:k * len(live)].reshape(k, len(live))
        expect = sorted(live, key=lambda o: (o - pr - 1) % world)
        assert (first == np.array(expect)).all()
