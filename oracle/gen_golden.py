"""Generate golden fixtures by executing the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE.  Run in the build container only (the reference cannot travel):

    python oracle/gen_golden.py [case ...]   # rewrites tests/golden/*.npz (all cases, or the named ones)

The reference's own classes are imported from /root/reference with the two import stubs of
oracle/stubs/ (torch_sparse -> torch.sparse.mm, dgl -> unused) and fed
seeded synthetic search logs written in the reference's on-disk format by
ihgnn_b200.synth.write_reference_files, i.e. through the reference's own
`GraphDataset.__init__` parsing and `PpsHyperGraph.from_interactions` Python loop.
Each fixture stores inputs, the reference `state_dict`, and the reference's outputs and
parameter gradients in fp32 and -- same modules cast to double -- fp64 (the arbiter).
Nothing from the reference's sources is copied; only numbers it computed are stored.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import tempfile

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(REPO, "oracle", "stubs"))
sys.path.insert(0, REFERENCE)
sys.path.insert(0, REPO)

from ihgnn_b200 import synth  # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    from Helpers.GlobalSettings import Gs, Gsv  # noqa: E402
    Gs.graph_completeness = Gsv.graph_uqi
    from Helpers.IOHelper import IOHelper  # noqa: E402
    IOHelper.warned_about_cannot_log = True
    from Helpers.Graph import PpsHyperGraph  # noqa: E402
    from Dataset import GraphDataset  # noqa: E402
    from Models import RawGnn, IHGNNLayer, HGCNLayer, GCNLayer, HemPredictionLayer  # noqa: E402
    from Helpers.Metrics import Metrics  # noqa: E402

CPU = torch.device("cpu")

CASES = {
    # name: synthetic log + model config.  Tiny on purpose: each npz stays well under 1 MB.
    "ihgnn_o3_L2_amazon": dict(U=60, Q=25, I=90, V=40, E=400, shape="amazon", seed=11,
                               gnn="IHGNN", L=2, order=3, d=16, batch=24),
    "ihgnn_o2_L1_cikm": dict(U=30, Q=12, I=50, V=25, E=220, shape="cikm", seed=12,
                             gnn="IHGNN", L=1, order=2, d=8, batch=16),
    "ihgnn_o1_L3_cikm": dict(U=45, Q=20, I=70, V=30, E=300, shape="cikm", seed=13,
                             gnn="IHGNN", L=3, order=1, d=32, batch=20),
    "gcn_L2_cikm": dict(U=40, Q=18, I=60, V=30, E=320, shape="cikm", seed=16,
                        gnn="GCN", L=2, order=1, d=16, batch=20),
    # Gs.Prediction.use_cosine_similarity = True (PredictionLayers.py:38-40)
    "ihgnn_o3_L1_cosine": dict(U=35, Q=14, I=55, V=25, E=260, shape="amazon", seed=17,
                               gnn="IHGNN", L=1, order=3, d=16, batch=18, cosine=True),
    # Gs.Query.transform = activation (EmbeddingLayers.py:40-44, :83-84), activation = the reference's nn.ReLU
    "ihgnn_o1_L2_qact": dict(U=32, Q=15, I=48, V=22, E=240, shape="cikm", seed=18,
                             gnn="IHGNN", L=2, order=1, d=32, batch=16, query_transform="activation"),
    "hgcn_L2_amazon": dict(U=50, Q=20, I=80, V=30, E=350, shape="amazon", seed=14,
                           gnn="HGCN", L=2, order=1, d=16, batch=20),
    # d=64 / 3 layers at the smallest size that still has heavy (Zipf head) nodes
    "ihgnn_o3_L3_d64": dict(U=100, Q=30, I=90, V=40, E=1200, shape="cikm", seed=15,
                            gnn="IHGNN", L=3, order=3, d=64, batch=33),
    # the cikm model shape (BASELINE.json configs[2]: d=128, 3 layers, order 3) on a small CIKM-shaped log
    "ihgnn_o3_L3_d128": dict(U=110, Q=36, I=84, V=36, E=1000, shape="cikm", seed=19,
                             gnn="IHGNN", L=3, order=3, d=128, batch=30),
}


def _np(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


def _cast_model_to_double(model):
    """fp64 arbiter: parameters AND the non-parameter graph tensors
    (Dv_neg_1, incidence) are cast, not recomputed."""
    model.double()
    for gnn in model.gnns:
        if hasattr(gnn, "Dv_neg_1"):
            gnn.Dv_neg_1 = gnn.Dv_neg_1.double()
        if hasattr(gnn, "Dv_neg_1_slash_2"):
            gnn.Dv_neg_1_slash_2 = gnn.Dv_neg_1_slash_2.double()
        if hasattr(gnn, "De_neg_1"):
            gnn.De_neg_1 = gnn.De_neg_1.double()
        for name in ("incidence", "incidence_t", "adjacency"):
            if hasattr(gnn, name):
                setattr(gnn, name, getattr(gnn, name).to(torch.float64))
    return model


def _layer_outputs(model):
    """Per-layer outputs, computed the way RawGnn.forward does (RawGnn.py:112-118)."""
    x = torch.cat(model.embeddings(None, None, None))
    outs = [x]
    h = x
    for gnn in model.gnns:
        h = gnn(h)
        outs.append(h)
    return outs


def _run(model, users, queries, items, flags, prefix, out, pos_log):
    model.zero_grad()
    scores = model(users, queries, items)
    loss = torch.nn.BCEWithLogitsLoss()(scores, flags.to(scores.dtype))   # Main.py:191
    loss.backward()
    out[f"{prefix}.scores"] = _np(scores)
    out[f"{prefix}.loss"] = _np(loss)
    for k, p in model.named_parameters():
        out[f"{prefix}.grad.{k}"] = _np(p.grad)
    with torch.no_grad():
        for li, o in enumerate(_layer_outputs(model)):
            out[f"{prefix}.layer_out.{li}"] = _np(o)
        # evaluation form (RawGnn.py:124-133, TestSearchLogDataLoader Dataset.py:324-329):
        # one (u, q) against ALL items, item_indices None
        model.save_features_for_test()
        I = model.dataset.item_count
        u0, q0 = int(users[0]), int(queries[0])
        ev = model(u0 * torch.ones(I, dtype=torch.long), q0 * torch.ones(I, dtype=torch.long), None)
        out[f"{prefix}.eval_scores"] = _np(ev)
        # the evaluation loop proper (TrainTestHelper.py:53-63): per search, all items scored, the
        # reference's own torch.sort top-10 and its own Metrics.calculate_on_all_items
        n_rank = min(8, int(users.numel()) // 11)
        top, top_s, mets, inter = [], [], [], []
        for b in range(n_rank):
            ub, qb = int(users[b]), int(queries[b])
            outb = model(ub * torch.ones(I, dtype=torch.long), qb * torch.ones(I, dtype=torch.long), None)
            _, idx = torch.sort(outb, descending=True)                         # Metrics.py:60
            top.append(_np(idx[:10]))
            top_s.append(_np(outb[idx[:10]]))
            sel = (pos_log[0] == ub) & (pos_log[1] == qb)
            items_b = sorted(set(int(x) for x in pos_log[2][sel]))             # get_interacted_items of that search
            m = Metrics.calculate_on_all_items(outb, items_b, None, True)
            mets.append([m.HitRatio_at10, m.NDCG_at10, m.MAP_at10])
            inter.append(items_b + [-1] * (16 - len(items_b)))
        out[f"{prefix}.rank_top10"] = np.stack(top).astype(np.int64)
        out[f"{prefix}.rank_scores"] = np.stack(top_s)
        out[f"{prefix}.rank_metrics"] = np.asarray(mets, dtype=np.float64)
        out["rank.interacted"] = np.asarray(inter, dtype=np.int64)
        model.clear_saved_feature()
    # conv-only metric M1: loss = sum(cat(outs,1)) w.r.t. X
    x = torch.cat(model.embeddings(None, None, None)).detach().clone().requires_grad_(True)
    outs = [x]
    h = x
    for gnn in model.gnns:
        h = gnn(h)
        outs.append(h)
    model.zero_grad()
    torch.cat(outs, 1).sum().backward()
    out[f"{prefix}.conv_dx"] = _np(x.grad)
    for k, p in model.named_parameters():
        if k.startswith("gnn_"):
            out[f"{prefix}.conv_grad.{k}"] = _np(p.grad)


def make_case(name: str, cfg: dict, outdir: str) -> None:
    Gs.Prediction.use_cosine_similarity = bool(cfg.get("cosine", False))
    Gs.Query.transform = Gsv.activation if cfg.get("query_transform") == "activation" else Gsv.mean
    try:
        _make_case(name, cfg, outdir)
    finally:
        Gs.Prediction.use_cosine_similarity = False
        Gs.Query.transform = Gsv.mean


def _make_case(name: str, cfg: dict, outdir: str) -> None:
    log = synth.make_search_log(cfg["U"], cfg["Q"], cfg["I"], cfg["E"], cfg["V"],
                                shape=cfg["shape"], seed=cfg["seed"], with_negatives=True)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        synth.write_reference_files(log, tmp)
        with contextlib.redirect_stdout(io.StringIO()):
            ds = GraphDataset(os.path.join(tmp, "graph_info.txt"),
                              os.path.join(tmp, "queries_multihot.txt"),
                              os.path.join(tmp, "train_data.csv"),
                              PpsHyperGraph, 10, 0, CPU)
    # the reference's own parse must reproduce the generator's positives, in order
    uqi = np.array([p.uqif()[:3] for p in ds.pos_interactions], dtype=np.int64)
    assert np.array_equal(uqi[:, 0], log.pos_user) and np.array_equal(uqi[:, 1], log.pos_query) \
        and np.array_equal(uqi[:, 2], log.pos_item), "generator/reference parse mismatch"

    g = ds.hypergraph            # the reference's Python-loop builder, Graph.py:94-134
    adj = g.Adjacency
    csr = adj.to_sparse_csr()
    out["counts"] = np.array([cfg["U"], cfg["Q"], cfg["I"], cfg["V"], cfg["E"]], dtype=np.int64)
    out["pos_user"], out["pos_query"], out["pos_item"] = log.pos_user, log.pos_query, log.pos_item
    out["bag_words"] = _np(ds.queries_for_embeddingbag)
    out["bag_offsets"] = _np(ds.queries_offset_for_embeddingbag)
    out["graph.I3"] = _np(g.I3)
    out["graph.coo_indices"] = _np(adj.indices())
    out["graph.coo_values"] = _np(adj.values())
    out["graph.crow"] = _np(csr.crow_indices())
    out["graph.col"] = _np(csr.col_indices())
    out["graph.VertexDegrees"] = _np(g.VertexDegrees)
    out["graph.EdgeDegrees"] = _np(g.EdgeDegrees)
    out["graph.EdgeCount"] = np.array(g.EdgeCount, dtype=np.int64)

    layer_type = {"IHGNN": IHGNNLayer, "HGCN": HGCNLayer, "GCN": GCNLayer}[cfg["gnn"]]
    if cfg["gnn"] == "GCN":
        g2 = ds.graph2d          # the reference's Pps2DGraph.from_interactions, Graph.py:19-81 (no self connection)
        out["graph2d.coo_indices"] = _np(g2.Adjacency.indices())
        out["graph2d.coo_values"] = _np(g2.Adjacency.values())
        out["graph2d.VertexDegrees"] = _np(g2.VertexDegrees)
    torch.manual_seed(1000 + cfg["seed"])
    model = RawGnn(device=CPU, dataset=ds, embedding_size=cfg["d"], gnn_layer_type=layer_type,
                   gnn_layer_count=cfg["L"], feature_interaction_order=cfg["order"],
                   phase2_attention=False, predictions=HemPredictionLayer, lambda_muq=0.5)
    for k, v in model.state_dict().items():
        out[f"state.{k}"] = _np(v)
    out["cfg.gnn"] = np.array(cfg["gnn"])
    out["cfg.L"] = np.array(cfg["L"])
    out["cfg.order"] = np.array(cfg["order"])
    out["cfg.d"] = np.array(cfg["d"])
    out["cfg.lambda_muq"] = np.array(0.5)
    out["cfg.cosine"] = np.array(bool(cfg.get("cosine", False)))
    act = ""
    if cfg.get("query_transform") == "activation":
        act = type(model.embeddings.query_transform[1]).__name__.lower()       # 'relu' (GlobalSettings.py:76)
    out["cfg.query_activation"] = np.array(act)

    # a training batch in TrainTestHelper.py:126-129 form: positives then 10 negatives each
    rng = np.random.default_rng(cfg["seed"] + 500)
    b = cfg["batch"]
    pick = rng.choice(cfg["E"], size=b, replace=False)
    pu, pq, pi = log.pos_user[pick], log.pos_query[pick], log.pos_item[pick]
    nu, nq = np.repeat(pu, 10), np.repeat(pq, 10)
    ni = rng.integers(0, cfg["I"], size=b * 10)
    users = torch.from_numpy(np.concatenate([pu, nu]))
    queries = torch.from_numpy(np.concatenate([pq, nq]))
    items = torch.from_numpy(np.concatenate([pi, ni]))
    flags = torch.cat([torch.ones(b), torch.zeros(b * 10)])
    out["batch.users"], out["batch.queries"], out["batch.items"] = _np(users), _np(queries), _np(items)
    out["batch.flags"] = _np(flags)

    pos_log = (log.pos_user, log.pos_query, log.pos_item)
    _run(model, users, queries, items, flags, "ref32", out, pos_log)

    # indexed embedding lookups, the Srrl client form (Srrl.py:74-94)
    with torch.no_grad():
        qi = torch.from_numpy(rng.integers(0, cfg["Q"], size=9))
        out["embed.query_indices"] = _np(qi)
        out["ref32.embed_user_idx"] = _np(model.embeddings.embed_user(users[:9]))
        out["ref32.embed_item_idx"] = _np(model.embeddings.embed_item(items[:9]))
        out["ref32.embed_query_idx"] = _np(model.embeddings.embed_query(qi))

    _cast_model_to_double(model)
    _run(model, users, queries, items, flags, "ref64", out, pos_log)

    path = os.path.join(outdir, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, E={g.EdgeCount}, "
          f"zero-degree nodes={(np.diff(out['graph.crow']) == 0).sum()}, "
          f"max degree={np.diff(out['graph.crow']).max()}")


def make_graph2d_variants(outdir):
    """Every branch of the reference's Pps2DGraph.from_interactions (Helpers/Graph.py:19-81): the four
    graph_completeness values x with / without self connections, on interactions whose flags run over
    {1, 2, 3} (a flag above 1 weights the u-i pair of graph_uqi, :44) with repeated (u, q, i) triples and
    isolated nodes; plus the reference GCNLayer (GnnLayers.py:9-45) forward and input / weight gradients on
    each graph, fp32 and fp64.  -> tests/golden/graph2d/variants.npz"""
    from Helpers.SearchLog import PosInteraction
    from Helpers.Graph import Pps2DGraph
    U, Q, I, E, d_in, d_out = 23, 9, 31, 140, 16, 16
    rng = np.random.default_rng(41)
    u = rng.integers(0, U - 2, size=E)                     # the last users / queries / items stay isolated
    q = rng.integers(0, Q - 1, size=E)
    i = rng.integers(0, I - 3, size=E)
    u[5:9], q[5:9], i[5:9] = u[4], q[4], i[4]              # repeated triples: coalesce() sums them
    fl = rng.integers(1, 4, size=E)
    inter = [PosInteraction(int(a), int(b), "", int(c), 1, 1, int(f), "") for a, b, c, f in zip(u, q, i, fl)]
    N = U + Q + I
    out = {"counts": np.array([U, Q, I, E]), "user": u, "query": q, "item": i, "flags": fl}
    torch.manual_seed(41)
    x = torch.randn(N, d_in)
    w = torch.randn(N, d_out)
    out["x"], out["w"] = _np(x), _np(w)

    class _DS:
        pass
    for comp in (Gsv.graph_uqi, Gsv.graph_only_uq, Gsv.graph_only_ui, Gsv.graph_only_qi):
        for self_conn in (False, True):
            Gs.graph_completeness = comp
            g2 = Pps2DGraph.from_interactions(inter, N, U, Q, self_conn, CPU)
            key = f"{comp}.{'self' if self_conn else 'noself'}"
            out[f"{key}.coo_indices"] = _np(g2.Adjacency.indices())
            out[f"{key}.coo_values"] = _np(g2.Adjacency.values())
            out[f"{key}.VertexDegrees"] = _np(g2.VertexDegrees)
            ds = _DS()
            ds.graph2d = g2
            torch.manual_seed(7)
            layer = GCNLayer(CPU, ds, d_in, d_out)
            if "lin.weight" not in out:
                out["lin.weight"], out["lin.bias"] = _np(layer.feature_transform.weight), _np(layer.feature_transform.bias)
            for prefix, dt in (("ref32", torch.float32), ("ref64", torch.float64)):
                if dt == torch.float64:
                    layer.double()
                    layer.Dv_neg_1_slash_2 = layer.Dv_neg_1_slash_2.double()
                    layer.adjacency = layer.adjacency.to(torch.float64)
                xx = x.to(dt).clone().requires_grad_(True)
                layer.zero_grad()
                y = layer(xx)
                (y * w.to(dt)).sum().backward()
                out[f"{key}.{prefix}.out"] = _np(y)
                out[f"{key}.{prefix}.dx"] = _np(xx.grad)
                out[f"{key}.{prefix}.dw"] = _np(layer.feature_transform.weight.grad)
    Gs.graph_completeness = Gsv.graph_uqi
    os.makedirs(os.path.join(outdir, "graph2d"), exist_ok=True)
    path = os.path.join(outdir, "graph2d", "variants.npz")
    np.savez_compressed(path, **out)
    print(f"graph2d_variants: {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    outdir = os.path.join(REPO, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    only = set(sys.argv[1:])                      # optional: names of the cases to (re)generate
    for name, cfg in CASES.items():
        if not only or name in only:
            make_case(name, cfg, outdir)
    if not only or "graph2d_variants" in only:
        make_graph2d_variants(outdir)


if __name__ == "__main__":
    main()
