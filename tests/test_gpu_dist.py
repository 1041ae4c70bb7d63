"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): the sharded IHGNN stack over
NCCL must reproduce the single-GPU stack on the same global hypergraph -- outputs, input
gradients and all-reduced weight gradients -- and stay bitwise deterministic."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import REPO

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, dim: int, routed: str, ret):
    sys.path.insert(0, REPO)
    os.environ["IHG_ROUTED_REDUCE"] = routed
    import torch.distributed as dist
    from ihgnn_b200 import synth
    from ihgnn_b200.dataset import GraphDataset
    from ihgnn_b200.dist import PartitionPlan, ShardedHyperGraph, ShardedIHGNNLayer, allreduce_dense_grads
    from ihgnn_b200.layers import IHGNNLayer
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        U, Q, I, E = 3000, 100, 1200, 40_000
        log = synth.make_search_log(U, Q, I, E, 50, shape="cikm", seed=9, zipf=1.0)
        gen = torch.Generator().manual_seed(11)
        x = (torch.randn(U + Q + I, dim, generator=gen) * 0.5)
        orders = [3, 1]
        # ---- single-GPU stack on the global graph (every rank computes it: it is the reference)
        ds = GraphDataset.from_search_log(log, dev)
        torch.manual_seed(5)
        ref_layers = [IHGNNLayer(dev, ds, dim, dim, o, False).to(dev) for o in orders]
        xg = x.to(dev).requires_grad_(True)
        h, outs_g = xg, []
        for L in ref_layers:
            h = L(h)
            outs_g.append(h)
        torch.cat(outs_g, 1).sum().backward()
        # ---- sharded stack with the same weights
        plan = PartitionPlan(log.pos_user, log.pos_query, log.pos_item, U, Q, I, world, rank, device=dev)   # planned on the GPU
        host_plan = PartitionPlan(log.pos_user, log.pos_query, log.pos_item, U, Q, I, world, rank)          # ... equals the CPU plan
        for name in PartitionPlan._TENSORS:
            assert np.array_equal(getattr(plan, name), getattr(host_plan, name)), name
        sg = ShardedHyperGraph(plan, dev)
        sh_layers = []
        for L, o in zip(ref_layers, orders):
            S = ShardedIHGNNLayer(sg, dim, o).to(dev)
            S.load_state_dict(L.state_dict(), strict=True)
            sh_layers.append(S)
        own = torch.from_numpy(plan.own_global_ids()).to(dev)

        def run():
            xo = x.to(dev)[own].clone().requires_grad_(True)
            for S in sh_layers:
                S.zero_grad()
            h, outs = xo, []
            for S in sh_layers:
                h = S(h)
                outs.append(h)
            torch.cat(outs, 1).sum().backward()
            for S in sh_layers:
                allreduce_dense_grads(S)
            return xo, outs

        xo, outs_s = run()
        err = 0.0
        for a, b in zip(outs_s, outs_g):
            err = max(err, float((a - b[own]).abs().max() / b.abs().max()))
        err = max(err, float((xo.grad - xg.grad[own]).abs().max() / xg.grad.abs().max()))
        for S, L in zip(sh_layers, ref_layers):
            for (k, ps), (_, pl) in zip(S.named_parameters(), L.named_parameters()):
                err = max(err, float((ps.grad - pl.grad).abs().max() / pl.grad.abs().max()))
        xo2, outs2 = run()
        same = all(torch.equal(a, b) for a, b in zip(outs_s, outs2)) and torch.equal(xo.grad, xo2.grad)
        ret[rank] = (err, bool(same))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("routed", ["1", "0"], ids=["routed-reduce", "pull-reduce"])
@pytest.mark.parametrize("dim", [64, 16])
def test_sharded_stack_matches_single_gpu(dim, routed):
    """routed: the reduction kernels write the halo partial sums straight into their owners' receive buffers
    (`ihg_*_routed`); pull: the owners read them out of the holders' tables (`ihg_halo_reduce`)."""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, dim, routed, ret), nprocs=world, join=True)
    assert len(ret) == world
    errs = [v[0] for v in ret.values()]
    assert max(errs) < 5e-6, dict(ret)       # different (but fixed) summation order across ranks
    assert all(v[1] for v in ret.values()), dict(ret)


def _rank_worker(rank: int, world: int, port: int, ret):
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    from ihgnn_b200 import functional as F_
    from ihgnn_b200 import synth
    from ihgnn_b200.dist import PartitionPlan, ShardedHyperGraph, ShardedRawGnn
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        U, Q, I, E, V, d, L = 3000, 100, 1200, 40_000, 50, 64, 2
        log = synth.make_search_log(U, Q, I, E, V, shape="cikm", seed=9, zipf=1.0)
        plan = PartitionPlan(log.pos_user, log.pos_query, log.pos_item, U, Q, I, world, rank)
        sg = ShardedHyperGraph(plan, dev)
        words, offsets = log.bag_inputs()
        torch.manual_seed(3)                                     # same replicated weights on every rank
        model = ShardedRawGnn(sg, words, offsets, V, d, L, 3).to(dev)
        feat = model.gather_features()
        own = torch.from_numpy(plan.own_global_ids()).to(dev)
        ok_layout = torch.equal(feat[own], model.output_features())
        # every rank holds the same replicated table
        chk = feat.double().sum().view(1)
        lst = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(lst, chk)
        ok_same = all(float(x) == float(lst[0]) for x in lst)
        gen = torch.Generator().manual_seed(17)
        B, C = 301, 200
        users = torch.randint(0, U, (B,), generator=gen).to(dev)
        queries = torch.randint(0, Q, (B,), generator=gen).to(dev)
        cand = torch.stack([torch.randperm(I, generator=gen)[:C] for _ in range(B)]).to(dev)
        share, ids, vals = model.rank(users, queries, cand, 10, features=feat)
        pl = model.prediction_layer
        ids_all, vals_all = F_.rank_topk(feat, users, queries, pl.items_bias, pl.lambda_muq, query_row0=U,
                                         item_row0=U + Q, item_count=I, candidates=cand, k=10)
        ok_rank = torch.equal(ids, ids_all[share]) and torch.equal(vals, vals_all[share])
        ok_share = share.tolist() == list(range(rank, B, world))
        # float64 check of the scores on the replicated table
        b0 = int(share[0])
        m = 0.5 * feat[queries[b0] + U].double() + 0.5 * feat[users[b0]].double()
        sc = (feat[cand[b0] + U + Q].double() * m).sum(1) + pl.items_bias[cand[b0]].double()
        top = torch.sort(sc, descending=True)[1][:10]
        ok_ref = cand[b0][top].tolist() == ids[0].tolist()
        ret[rank] = (bool(ok_layout), bool(ok_same), bool(ok_rank), bool(ok_share), bool(ok_ref))
    finally:
        dist.destroy_process_group()


def test_sharded_ranking_replicated_features():
    """BASELINE.json configs[4] layout: features replicated by one all-gather, searches split across
    the ranks, no per-query communication; each rank's results equal the single-device ranking."""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_rank_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    assert all(all(v) for v in ret.values()), dict(ret)


def _train_worker(rank: int, world: int, port: int, ret):
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    from ihgnn_b200 import HemPredictionLayer, IHGNNLayer, RawGnn, synth
    from ihgnn_b200.dataset import GraphDataset
    from ihgnn_b200.dist import PartitionPlan, ShardedHyperGraph, ShardedRawGnn
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        U, Q, I, E, V, d, L = 3000, 100, 1200, 40_000, 50, 64, 2
        log = synth.make_search_log(U, Q, I, E, V, shape="cikm", seed=9, zipf=1.0)
        # ---- the single-GPU model on the global graph: identical on every rank (same seed)
        ds = GraphDataset.from_search_log(log, dev)
        torch.manual_seed(5)
        ref = RawGnn(device=dev, dataset=ds, embedding_size=d, gnn_layer_type=IHGNNLayer, gnn_layer_count=L,
                     feature_interaction_order=3, phase2_attention=False, predictions=HemPredictionLayer,
                     lambda_muq=0.5).to(dev)
        # ---- the sharded model with the same parameters
        plan = PartitionPlan(log.pos_user, log.pos_query, log.pos_item, U, Q, I, world, rank)
        sg = ShardedHyperGraph(plan, dev)
        words, offsets = log.bag_inputs()
        sh = ShardedRawGnn(sg, words, offsets, V, d, L, 3).to(dev)
        u0, u1, i0, i1 = int(plan.ub[rank]), int(plan.ub[rank + 1]), int(plan.ib[rank]), int(plan.ib[rank + 1])
        with torch.no_grad():
            sh.embedding_user.copy_(ref.embeddings.embedding_user.weight[1:][u0:u1])
            sh.embedding_item.copy_(ref.embeddings.embedding_item.weight[1:][i0:i1])
            sh.embedding_bag_vocabulary.copy_(ref.embeddings.embedding_bag_vocabulary.weight)
            sh.prediction_layer.items_bias.copy_(ref.prediction_layer.items_bias)
            for a, b in zip(sh.gnns, ref.gnns):
                a.load_state_dict(b.state_dict(), strict=True)
        rng = np.random.default_rng(3)
        B = 64
        pick = rng.choice(E, size=B, replace=False)
        users = torch.from_numpy(np.concatenate([log.pos_user[pick], np.repeat(log.pos_user[pick], 10)])).to(dev)
        queries = torch.from_numpy(np.concatenate([log.pos_query[pick], np.repeat(log.pos_query[pick], 10)])).to(dev)
        items = torch.from_numpy(np.concatenate([log.pos_item[pick], rng.integers(0, I, size=10 * B)])).to(dev)
        flags = torch.cat([torch.ones(B), torch.zeros(10 * B)]).to(dev)

        def step(m, sync=None):
            m.zero_grad(set_to_none=True)
            sc = m(users, queries, items)
            loss = torch.nn.functional.binary_cross_entropy_with_logits(sc, flags)
            loss.backward()
            if sync:
                sync()
            return sc.detach(), float(loss)

        sc_r, loss_r = step(ref)
        sc_s, loss_s = step(sh, sh.sync_grads)
        rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
        errs = {"scores": rel(sc_s, sc_r), "loss": abs(loss_s - loss_r) / abs(loss_r),
                "user": rel(sh.embedding_user.grad, ref.embeddings.embedding_user.weight.grad[1:][u0:u1]),
                "item": rel(sh.embedding_item.grad, ref.embeddings.embedding_item.weight.grad[1:][i0:i1]),
                "vocab": rel(sh.embedding_bag_vocabulary.grad, ref.embeddings.embedding_bag_vocabulary.weight.grad),
                "bias": rel(sh.prediction_layer.items_bias.grad, ref.prediction_layer.items_bias.grad)}
        for k, (a, b) in enumerate(zip(sh.gnns, ref.gnns)):
            for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
                errs[f"gnn{k}.{n}"] = rel(pa.grad, pb.grad)
        sc_s2, loss_s2 = step(sh, sh.sync_grads)
        ret[rank] = (errs, bool(torch.equal(sc_s, sc_s2)) and loss_s == loss_s2)
    finally:
        dist.destroy_process_group()


def test_sharded_training_step_matches_single_gpu():
    """ShardedRawGnn.forward -> BCE -> backward -> sync_grads (what bench.py's e2e leg runs at N > 1)
    against the single-GPU RawGnn with the same parameters: scores, loss, the sharded embedding
    gradients and the all-reduced replicated gradients; deterministic."""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_train_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r, (errs, same) in ret.items():
        bad = {k: v for k, v in errs.items() if not v < 1e-5}
        assert not bad, (r, bad)
        assert same, r
