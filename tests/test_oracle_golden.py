"""Pins the oracle (oracle/ihgnn_oracle.py) against outputs of the reference itself.

The reference has no tests or golden vectors, so the fixtures under
tests/golden/ were produced by oracle/gen_golden.py executing the unmodified reference
classes on CPU.  Indices must agree bit-exactly; fp32 values must agree far inside the 1e-5
parity budget (same ATen ops in the same order), and the fp64 arbiter to ~1e-12.
"""
import numpy as np
import pytest
import torch

from helpers import batch_of, max_rel, oracle_graph, oracle_model, orc


def test_graph_indices_bit_exact(golden):
    g = oracle_graph(golden)
    assert np.array_equal(g.I3.numpy(), golden["graph.I3"])
    assert np.array_equal(g.rowptr.numpy(), golden["graph.crow"])
    assert np.array_equal(g.col.numpy(), golden["graph.col"])
    assert np.array_equal(torch.stack([g.row, g.col]).numpy(), golden["graph.coo_indices"])
    assert np.array_equal(g.VertexDegrees.numpy(), golden["graph.VertexDegrees"])
    assert np.array_equal(g.EdgeDegrees.numpy(), golden["graph.EdgeDegrees"])
    assert g.EdgeCount == int(golden["graph.EdgeCount"])
    assert np.all(golden["graph.coo_values"] == 1.0)


def _check(golden, dtype, prefix, tol):
    m = oracle_model(golden, dtype)
    users, queries, items, flags = batch_of(golden)
    scores = m.forward(users, queries, items)
    loss = orc.bce_with_logits_mean(scores, flags.to(dtype))
    loss.backward()
    assert max_rel(scores.detach().numpy(), golden[f"{prefix}.scores"]) <= tol
    assert max_rel(loss.detach().numpy(), golden[f"{prefix}.loss"]) <= tol
    for k, gr in m.grads().items():
        assert max_rel(gr.numpy(), golden[f"{prefix}.grad.{k}"]) <= tol, k
    with torch.no_grad():
        outs = m.conv_stack(m.input_features())
        for li, o in enumerate(outs):
            assert max_rel(o.numpy(), golden[f"{prefix}.layer_out.{li}"]) <= tol, li
        I = m.I
        ev = m.forward(users[0] * torch.ones(I, dtype=torch.long),
                       queries[0] * torch.ones(I, dtype=torch.long), None)
        assert max_rel(ev.numpy(), golden[f"{prefix}.eval_scores"]) <= tol
    x = m.input_features().detach()
    dx = orc.conv_fwd_bwd(m, x)
    assert max_rel(dx.numpy(), golden[f"{prefix}.conv_dx"]) <= tol
    for k, gr in m.grads().items():
        if k.startswith("gnn_"):
            assert max_rel(gr.numpy(), golden[f"{prefix}.conv_grad.{k}"]) <= tol, k


def test_model_fp32_matches_reference(golden):
    # same ops, same order: expect ~0; allow a few ulp for MKL threading differences
    _check(golden, torch.float32, "ref32", 2e-6)


def test_model_fp64_matches_reference(golden):
    _check(golden, torch.float64, "ref64", 1e-12)


def test_reference_fp32_noise_floor_is_inside_budget(golden):
    """The reference's own fp32-vs-fp64 error bounds what 1e-5 parity can mean."""
    worst = 0.0
    for k in golden:
        if k.startswith("ref32.") and k != "ref32.loss" and "ref64." + k[len("ref32."):] in golden:
            worst = max(worst, max_rel(golden[k], golden["ref64." + k[len("ref32."):]]))
    assert worst < 1e-5, worst


def test_indexed_embedding_lookups(golden):
    m = oracle_model(golden)
    p = m.params
    users, queries, items, _ = batch_of(golden)
    with torch.no_grad():
        eu = orc.embed_user(p["embeddings.embedding_user.weight"], users[:9])
        ei = orc.embed_item(p["embeddings.embedding_item.weight"], items[:9])
        transform = None
        if m.query_activation:                          # Gs.Query.transform == activation fixtures
            transform = (p["embeddings.query_transform.0.weight"], p["embeddings.query_transform.0.bias"],
                         m.query_activation)
        eq = orc.embed_query(p["embeddings.embedding_bag_vocabulary.weight"], m.bag_words,
                             m.bag_offsets, torch.from_numpy(golden["embed.query_indices"]), transform)
    assert np.array_equal(eu.numpy(), golden["ref32.embed_user_idx"])
    assert np.array_equal(ei.numpy(), golden["ref32.embed_item_idx"])
    assert max_rel(eq.numpy(), golden["ref32.embed_query_idx"]) <= 1e-7


def test_ranking_matches_reference_eval_loop(golden):
    """oracle.rank_topk / metrics_at_10 against the reference's own evaluation loop
    (TrainTestHelper.py:53-63: all-item scores -> torch.sort top-10 -> Metrics.calculate_on_all_items)."""
    users, queries, _, _ = batch_of(golden)
    n = golden["ref64.rank_top10"].shape[0]
    m = oracle_model(golden, torch.float64)
    ids, vals = orc.rank_topk(m, users[:n], queries[:n], None, 10)
    assert np.array_equal(ids.numpy(), golden["ref64.rank_top10"])
    assert max_rel(vals.numpy(), golden["ref64.rank_scores"]) <= 1e-12
    for b in range(n):
        inter = [int(x) for x in golden["rank.interacted"][b] if x >= 0]
        got = orc.metrics_at_10(golden["ref64.rank_top10"][b], inter)
        assert np.allclose(got, golden["ref64.rank_metrics"][b], rtol=0, atol=1e-12), (b, got)
    # candidate-list form == all-items form restricted to the list
    rng = np.random.default_rng(3)
    cand = torch.from_numpy(np.stack([rng.permutation(m.I)[: m.I // 2] for _ in range(n)]))
    ids_c, vals_c = orc.rank_topk(m, users[:n], queries[:n], cand, 5)
    f = m.features()
    for b in range(n):
        full = m.forward(users[b] * torch.ones(m.I, dtype=torch.long), queries[b] * torch.ones(m.I, dtype=torch.long),
                         None, features=f).detach()
        want = sorted(cand[b].tolist(), key=lambda i: -float(full[i]))[:5]
        assert ids_c[b].tolist() == want


def test_graph2d_bit_exact(golden):
    """oracle.build_graph2d against the reference's Pps2DGraph.from_interactions (Graph.py:19-81)."""
    import pytest
    if "graph2d.coo_indices" not in golden:
        pytest.skip("fixture without the 2-D graph")
    U, Q, I, V, E = (int(x) for x in golden["counts"])
    adj, deg = orc.build_graph2d(golden["pos_user"], golden["pos_query"], golden["pos_item"], U, Q, I, False)
    assert np.array_equal(adj.indices().numpy(), golden["graph2d.coo_indices"])
    assert np.array_equal(adj.values().numpy(), golden["graph2d.coo_values"])
    assert np.array_equal(deg.numpy(), golden["graph2d.VertexDegrees"])


GRAPH2D_VARIANTS = [(c, sc) for c in ("uqi", "uq", "ui", "qi") for sc in (False, True)]


def load_graph2d_variants():
    import os
    from conftest import GOLDEN_DIR
    with np.load(os.path.join(GOLDEN_DIR, "graph2d", "variants.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("completeness,self_conn", GRAPH2D_VARIANTS)
def test_graph2d_every_branch_bit_exact(completeness, self_conn):
    """oracle.build_graph2d against every branch of the reference's Pps2DGraph.from_interactions
    (Graph.py:40-65: four graph_completeness values, flags 1..3, with / without self connections) and
    oracle.gcn_layer against the reference GCNLayer on each graph (fixture: oracle/gen_golden.py
    make_graph2d_variants)."""
    z = load_graph2d_variants()
    U, Q, I, E = (int(v) for v in z["counts"])
    key = f"{completeness}.{'self' if self_conn else 'noself'}"
    adj, deg = orc.build_graph2d(z["user"], z["query"], z["item"], U, Q, I, self_conn, completeness, z["flags"])
    assert np.array_equal(adj.indices().numpy(), z[f"{key}.coo_indices"])
    assert np.array_equal(adj.values().numpy(), z[f"{key}.coo_values"])
    assert np.array_equal(deg.numpy(), z[f"{key}.VertexDegrees"])
    if completeness == "uqi":
        assert z[f"{key}.coo_values"].max() > 3                      # flag weights and repeated triples did add up
    x = torch.from_numpy(z["x"]).double().requires_grad_(True)
    w = torch.from_numpy(z["lin.weight"]).double().requires_grad_(True)
    b = torch.from_numpy(z["lin.bias"]).double()
    y = orc.gcn_layer(x, adj.double(), deg.pow(-0.5).double(), w, b)
    (y * torch.from_numpy(z["w"]).double()).sum().backward()
    assert max_rel(y.detach().numpy(), z[f"{key}.ref64.out"]) <= 1e-12
    assert max_rel(x.grad.numpy(), z[f"{key}.ref64.dx"]) <= 1e-12
    assert max_rel(w.grad.numpy(), z[f"{key}.ref64.dw"]) <= 1e-12
