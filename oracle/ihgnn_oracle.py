"""CPU oracle for the IHGNN hypergraph-convolution hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / `--impl reference` legs may import this module, and only as the
checker / the CPU baseline.  The product (ihgnn_b200/) never imports it and has no CPU path.

It is a restatement, in plain torch-CPU ops (the reference itself is pure PyTorch, so the
same ATen ops in the same order are the faithful CPU form), of:

  build_hypergraph     Helpers/Graph.py:94-134        PpsHyperGraph.from_interactions
  embed_all / embed_*  Models/EmbeddingLayers.py:63-91 EmbeddingLayer.forward / embed_*
  feature_interactor   Models/CommonLayers.py:58-87   FeatureInteractor.forward
  ihgnn_layer          Models/GnnLayers.py:221-236    IHGNNLayer.forward
  hgcn_layer           Models/GnnLayers.py:142-153    HGCNLayer.forward
  build_graph2d / gcn_layer  Helpers/Graph.py:19-81, Models/GnnLayers.py:9-45  Pps2DGraph, GCNLayer.forward
  hem_score            Models/PredictionLayers.py:21-44 HemPredictionLayer.forward
  rawgnn_features/forward  Models/RawGnn.py:104-144   RawGnn.forward
  rank_topk / metrics_at_10  Dataset.py:324-329 + Helpers/Metrics.py:47-110  the evaluation loop

The edge->node reduction of the reference goes through `torch_sparse.matmul`
(rusty1s/pytorch_sparse, NOT vendored and NOT version-pinned by the reference: it ships no
requirements file).  Its published semantics -- coalesce = sum duplicates, matmul =
sum-reduce CSR SpMM -- are restated with torch.sparse.mm, as SURVEY.md section 8c
prescribes and as BASELINE.json's north_star names ("reference torch.sparse.mm path").

PINNING: the reference has no tests or golden vectors of its own.
This oracle is pinned against outputs of the reference itself, executed in the build
container by oracle/gen_golden.py (fixtures under tests/golden/), by
tests/test_oracle_golden.py -- indices bit-exact, fp32 values to ~1 ulp-level agreement
-- and the host-side readers / drop-in wiring are checked against the live reference where
/root/reference is present (tests/test_loaders.py, tests/test_dropin_wiring.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch


# --------------------------------------------------------------------------------------
# graph / index construction  (Helpers/Graph.py:94-134)
# --------------------------------------------------------------------------------------
class OracleHyperGraph:
    """Index arrays of the 3-uniform hypergraph, int64 as in the reference."""
    I3: torch.Tensor              # [E,3] global node ids (u, q+U, i+U+Q)   Graph.py:110-117,129
    rowptr: torch.Tensor          # [N+1]  CSR of Adjacency.coalesce()      Graph.py:123-128
    col: torch.Tensor             # [3E]   edge ids, ascending inside a node
    row: torch.Tensor             # [3E]   node id of every incidence (COO row index)
    VertexDegrees: torch.Tensor   # fp32 [N,1], zero degree stored as 1e-8  Graph.py:112,120,131
    EdgeDegrees: torch.Tensor     # fp32 [E,1] == 3                         Graph.py:132
    node_count: int
    EdgeCount: int

    def adjacency(self, dtype=torch.float32) -> torch.Tensor:
        """The coalesced sparse COO incidence matrix [N,E] with unit values."""
        idx = torch.stack([self.row, self.col])
        vals = torch.ones(idx.shape[1], dtype=dtype)
        return torch.sparse_coo_tensor(idx, vals, (self.node_count, self.EdgeCount),
                                       is_coalesced=True)


def build_hypergraph(user: Sequence[int], query: Sequence[int], item: Sequence[int],
                     user_count: int, query_count: int, item_count: int) -> OracleHyperGraph:
    """Graph.py:94-134 without the Python per-edge loop: every positive (u,q,i) becomes one
    hyperedge numbered by its position; incidences are ordered as coalesce() orders them,
    i.e. lexicographically by (node, edge)."""
    u = np.asarray(user, dtype=np.int64)
    q = np.asarray(query, dtype=np.int64) + user_count                    # Graph.py:110
    i = np.asarray(item, dtype=np.int64) + user_count + query_count       # Graph.py:111
    E = u.shape[0]
    N = user_count + query_count + item_count
    i3 = np.stack([u, q, i], axis=1) if E else np.zeros((0, 3), dtype=np.int64)
    nodes = i3.reshape(-1)                                  # COO rows in insertion order
    edges = np.repeat(np.arange(E, dtype=np.int64), 3)      # Graph.py:114-116
    order = np.lexsort((edges, nodes))                      # coalesce(): sort by (row, col)
    row = nodes[order]
    col = edges[order]
    counts = np.bincount(nodes, minlength=N).astype(np.int64)
    rowptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    deg = counts.astype(np.float32)
    deg[deg == 0] = np.float32(1e-8)                        # Graph.py:120
    g = OracleHyperGraph()
    g.I3 = torch.from_numpy(i3)
    g.rowptr = torch.from_numpy(rowptr)
    g.col = torch.from_numpy(col)
    g.row = torch.from_numpy(row)
    g.VertexDegrees = torch.from_numpy(deg).view(-1, 1)
    g.EdgeDegrees = torch.full((E, 1), 3.0, dtype=torch.float32)
    g.node_count = N
    g.EdgeCount = E
    return g


# --------------------------------------------------------------------------------------
# embedding lookups  (Models/EmbeddingLayers.py:63-91)
# --------------------------------------------------------------------------------------
def embed_user(weight_user: torch.Tensor, user_indices: Optional[torch.Tensor] = None):
    """EmbeddingLayers.py:70-71: one-hot value == index + 1, row 0 is padding."""
    if user_indices is None:
        return weight_user[1:]
    return weight_user[user_indices + 1]


def embed_item(weight_item: torch.Tensor, item_indices: Optional[torch.Tensor] = None):
    """EmbeddingLayers.py:73-74."""
    if item_indices is None:
        return weight_item[1:]
    return weight_item[item_indices + 1]


def embed_query(weight_vocab: torch.Tensor, bag_words: torch.Tensor, bag_offsets: torch.Tensor,
                query_indices: Optional[torch.Tensor] = None, transform=None):
    """EmbeddingLayers.py:76-91: EmbeddingBag(mean) over all Q queries, then an optional row select
    (empty bags give zeros).  Gs.Query.transform == activation (:83-84, :40-44) applies
    `transform` = (Linear weight, Linear bias, activation name 'relu' | 'tanh') afterwards."""
    out = torch.nn.functional.embedding_bag(bag_words, weight_vocab, bag_offsets, mode="mean")
    if query_indices is not None:
        out = out[query_indices]
    if transform is not None:
        w, b, act = transform
        out = getattr(torch, act)(torch.nn.functional.linear(out, w, b))
    return out


def embed_all(weight_user, weight_vocab, weight_item, bag_words, bag_offsets, transform=None):
    """EmbeddingLayer.forward(None, None, None) followed by RawGnn's cat (RawGnn.py:112)."""
    return torch.cat([embed_user(weight_user),
                      embed_query(weight_vocab, bag_words, bag_offsets, None, transform),
                      embed_item(weight_item)])


# --------------------------------------------------------------------------------------
# conv layers
# --------------------------------------------------------------------------------------
def feature_interactor(node_features: torch.Tensor, i3: torch.Tensor,
                       agg_weight: torch.Tensor, agg_bias: torch.Tensor, order: int):
    """CommonLayers.py:58-87.  order 1: X'[I3] -> [E,3d] -> Linear; order 2/3: u,q,i rows,
    pairwise (and triple) Hadamard products, cat, Linear(K*d -> d), K = 3/6/7."""
    d = node_features.shape[1]
    if order == 1:
        sel = node_features[i3]                                   # :62
        cat = sel.reshape(-1, 3 * d)                              # :64
    else:
        u = node_features[i3[:, 0]]                               # :70-72
        q = node_features[i3[:, 1]]
        i = node_features[i3[:, 2]]
        uq = u * q                                                # :74-76
        qi = q * i
        iu = i * u
        if order == 3:
            uqi = uq * i                                          # :79
            cat = torch.cat([u, q, i, uq, qi, iu, uqi], 1)        # :84
        else:
            cat = torch.cat([u, q, i, uq, qi, iu], 1)             # :82
    return torch.nn.functional.linear(cat, agg_weight, agg_bias)  # :66 / :85


def ihgnn_layer(x: torch.Tensor, graph: OracleHyperGraph, adjacency: torch.Tensor,
                dv_neg_1: torch.Tensor, transform_weight, transform_bias,
                agg_weight, agg_bias, order: int):
    """GnnLayers.py:221-236 (phase2_attention False): Linear -> FeatureInteractor ->
    incidence SpMM -> multiply by the precomputed reciprocal degree (GnnLayers.py:187)."""
    xp = torch.nn.functional.linear(x, transform_weight, transform_bias)   # :224
    ef = feature_interactor(xp, graph.I3, agg_weight, agg_bias, order)      # :225
    s = torch.sparse.mm(adjacency, ef)                                      # :233
    return dv_neg_1 * s                                                     # :234


def hgcn_layer(x: torch.Tensor, adjacency: torch.Tensor, adjacency_t: torch.Tensor,
               dv_neg_half: torch.Tensor, de_neg_1: torch.Tensor,
               transform_weight, transform_bias):
    """GnnLayers.py:142-153: Dv^-1/2 H De^-1 H^T Dv^-1/2 Linear(X)."""
    h = torch.nn.functional.linear(x, transform_weight, transform_bias)    # :145
    h = dv_neg_half * h                                                     # :146
    e = torch.sparse.mm(adjacency_t, h)                                     # :148
    e = de_neg_1 * e                                                        # :149
    o = torch.sparse.mm(adjacency, e)                                       # :151
    return dv_neg_half * o                                                  # :152


def build_graph2d(user, query, item, user_count: int, query_count: int, item_count: int,
                  use_self_connection: bool = False, completeness: str = "uqi", flags=None):
    """Pps2DGraph.from_interactions (Helpers/Graph.py:19-81).  graph_uqi (:40-45): per interaction the six
    directed pairs u-q, q-i, i-u, i-q, q-u, u-i with values 1, 1, flag, 1, 1, flag, degrees += 2 per node;
    graph_only_uq / ui / qi (:46-63): the one pair both ways with value 1, degrees += 1 for its two nodes.
    Coalesced (duplicates summed, :71-77); degree 0 -> 1e-8 without self connections (:67-68).  `flags`
    None = all 1 (treat_all_1, Dataset.py:200).  Returns (coalesced COO [N,N] float32, VertexDegrees [N,1])."""
    u = torch.as_tensor(np.asarray(user, dtype=np.int64))
    q = torch.as_tensor(np.asarray(query, dtype=np.int64)) + user_count
    i = torch.as_tensor(np.asarray(item, dtype=np.int64)) + user_count + query_count
    n = user_count + query_count + item_count
    one = torch.ones(u.numel(), dtype=torch.float32)
    f = one if flags is None else torch.as_tensor(np.asarray(flags), dtype=torch.float32)
    if completeness == "uqi":
        rows = torch.stack([u, q, i, i, q, u], 1).reshape(-1)            # insertion order of :42
        cols = torch.stack([q, i, u, q, u, i], 1).reshape(-1)            # :43
        vals = torch.stack([one, one, f, one, one, f], 1).reshape(-1)    # :44
        touched, per_node = [u, q, i], 2.0                               # :45
    else:
        a, b = {"uq": (u, q), "ui": (u, i), "qi": (q, i)}[completeness]  # :46-63
        rows = torch.stack([a, b], 1).reshape(-1)
        cols = torch.stack([b, a], 1).reshape(-1)
        vals = torch.ones(rows.numel(), dtype=torch.float32)
        touched, per_node = [a, b], 1.0
    deg = torch.zeros(n, dtype=torch.float32)
    if use_self_connection:
        eye = torch.arange(n)
        rows, cols = torch.cat([eye, rows]), torch.cat([eye, cols])      # :28
        vals = torch.cat([torch.ones(n, dtype=torch.float32), vals])
        deg += 1                                                         # :29
    deg += per_node * torch.bincount(torch.cat(touched), minlength=n).to(torch.float32)
    if not use_self_connection:
        deg[deg == 0] = 1e-8                                             # :67-68
    adj = torch.sparse_coo_tensor(torch.stack([rows, cols]), vals, (n, n)).coalesce()
    return adj, deg.view(-1, 1)


def gcn_layer(x: torch.Tensor, adjacency2d: torch.Tensor, dv_neg_half: torch.Tensor,
              transform_weight, transform_bias):
    """GnnLayers.py:27-45: D^-1/2 A D^-1/2 applied after the Linear when d_in >= d_out (:33-38),
    before it otherwise (:39-43)."""
    d_out, d_in = transform_weight.shape
    if d_in >= d_out:
        h = torch.nn.functional.linear(x, transform_weight, transform_bias)    # :34
        h = dv_neg_half * h                                                     # :35
        h = torch.sparse.mm(adjacency2d, h)                                     # :36
        return dv_neg_half * h                                                  # :37
    h = dv_neg_half * x                                                         # :40
    h = torch.sparse.mm(adjacency2d, h)                                         # :41
    h = dv_neg_half * h                                                         # :42
    return torch.nn.functional.linear(h, transform_weight, transform_bias)      # :43


def hem_score(user_feature: Optional[torch.Tensor], query_feature: torch.Tensor,
              item_feature: torch.Tensor, items_bias: torch.Tensor,
              item_indices: Optional[torch.Tensor], lambda_muq: float, cosine: bool = False):
    """PredictionLayers.py:21-44: dot-product branch (:41-43) or, with
    Gs.Prediction.use_cosine_similarity, torch.cosine_similarity(item, m) + bias (:38-40)."""
    bias = items_bias if item_indices is None else items_bias[item_indices]   # :30-31
    if user_feature is not None:
        m = lambda_muq * query_feature + (1 - lambda_muq) * user_feature      # :35
    else:
        m = query_feature                                                     # :37
    if cosine:
        return torch.cosine_similarity(item_feature, m) + bias                # :39-40
    return (item_feature * m).sum(1) + bias                                   # :42-43


# --------------------------------------------------------------------------------------
# the RawGnn caller  (Models/RawGnn.py)
# --------------------------------------------------------------------------------------
class OracleModel:
    """Holds a `state_dict` with the reference's key names (checkpoint
    row) and evaluates RawGnn.forward on CPU.  dtype float32 reproduces the reference's fp32
    path; float64 is the arbiter used for tolerances."""

    def __init__(self, state: Dict[str, torch.Tensor], graph: OracleHyperGraph,
                 bag_words: torch.Tensor, bag_offsets: torch.Tensor,
                 user_count: int, query_count: int, item_count: int,
                 layer_type: str = "IHGNN", layer_count: int = 2, order: int = 3,
                 lambda_muq: float = 0.5, dtype=torch.float32, requires_grad: bool = True,
                 cosine: bool = False, query_activation: Optional[str] = None):
        self.dtype = dtype
        self.cosine = cosine
        self.query_activation = query_activation          # None (mean) | 'relu' | 'tanh'  (Gs.Query.transform)
        self.graph = graph
        self.layer_type = layer_type
        self.layer_count = layer_count
        self.order = order
        self.lambda_muq = lambda_muq
        self.U, self.Q, self.I = user_count, query_count, item_count
        self.bag_words, self.bag_offsets = bag_words, bag_offsets
        self.params = {k: torch.as_tensor(v).detach().clone().to(dtype).requires_grad_(requires_grad)
                       for k, v in state.items()}
        self.adjacency = graph.adjacency(dtype)
        self.adjacency_t = self.adjacency.t().coalesce()
        # reciprocals are taken in fp32 exactly like GnnLayers.py:187 / :133-134, then cast
        # (the fp64 arbiter of SURVEY.md section 8c casts Dv_neg_1 rather than recomputing it)
        self.dv_neg_1 = graph.VertexDegrees.pow(-1).to(dtype)
        self.dv_neg_half = graph.VertexDegrees.pow(-0.5).to(dtype)
        self.de_neg_1 = graph.EdgeDegrees.pow(-1).to(dtype)
        if layer_type == "GCN":
            i3 = graph.I3.numpy()
            adj2d, deg2d = build_graph2d(i3[:, 0], i3[:, 1] - user_count, i3[:, 2] - user_count - query_count,
                                         user_count, query_count, item_count, False)      # Dataset.py:87: no self connection
            self.adjacency2d = adj2d.to(dtype)
            self.dv2d_neg_half = deg2d.pow(-0.5).to(dtype)                                # GnnLayers.py:24 (fp32 pow, then cast)

    def layer_order(self, layer: int) -> int:
        """RawGnn.py:76-78: interactions above order 1 apply to layer 0 only."""
        return self.order if (layer == 0 or self.order == 1) else 1

    def input_features(self) -> torch.Tensor:
        p = self.params
        transform = None
        if self.query_activation is not None:
            transform = (p["embeddings.query_transform.0.weight"], p["embeddings.query_transform.0.bias"],
                         self.query_activation)
        return embed_all(p["embeddings.embedding_user.weight"],
                         p["embeddings.embedding_bag_vocabulary.weight"],
                         p["embeddings.embedding_item.weight"],
                         self.bag_words, self.bag_offsets, transform)

    def conv_stack(self, x: torch.Tensor) -> List[torch.Tensor]:
        """RawGnn.py:113-118: outputs of every layer, input first."""
        p = self.params
        outs = [x]
        h = x
        for k in range(self.layer_count):
            tw, tb = p[f"gnn_{k}.feature_transform.weight"], p[f"gnn_{k}.feature_transform.bias"]
            if self.layer_type == "IHGNN":
                h = ihgnn_layer(h, self.graph, self.adjacency, self.dv_neg_1, tw, tb,
                                p[f"gnn_{k}.feature_interactor.aggregation.weight"],
                                p[f"gnn_{k}.feature_interactor.aggregation.bias"],
                                self.layer_order(k))
            elif self.layer_type == "HGCN":
                h = hgcn_layer(h, self.adjacency, self.adjacency_t, self.dv_neg_half,
                               self.de_neg_1, tw, tb)
            elif self.layer_type == "GCN":
                h = gcn_layer(h, self.adjacency2d, self.dv2d_neg_half, tw, tb)
            else:
                raise NotImplementedError(self.layer_type)
            outs.append(h)
        return outs

    def features(self) -> torch.Tensor:
        """RawGnn.py:110-122: F = cat(all layer outputs, 1), shape [N, d(1+L)]."""
        return torch.cat(self.conv_stack(self.input_features()), 1)

    def forward(self, users: torch.Tensor, queries: torch.Tensor,
                items: Optional[torch.Tensor] = None, features: Optional[torch.Tensor] = None):
        """RawGnn.py:104-144."""
        f = self.features() if features is None else features
        fu = f[users]                                          # :128
        fq = f[queries + self.U]                               # :129
        if items is not None:
            fi = f[items + self.U + self.Q]                    # :131
        else:
            fi = f[self.U + self.Q:]                           # :133
        return hem_score(fu, fq, fi, self.params["prediction_layer.items_bias"], items,
                         self.lambda_muq, self.cosine)

    def grads(self) -> Dict[str, torch.Tensor]:
        return {k: (v.grad if v.grad is not None else torch.zeros_like(v))
                for k, v in self.params.items()}


def rank_topk(model: "OracleModel", users: torch.Tensor, queries: torch.Tensor,
              candidates: Optional[torch.Tensor] = None, k: int = 10,
              features: Optional[torch.Tensor] = None):
    """The reference's evaluation of one search per loop iteration, restated per query:
    TestSearchLogDataLoader.__iter__ (Dataset.py:324-329: users = u * ones(I), queries = q * ones(I))
    -> RawGnn.forward(users, queries, None) on the saved features (RawGnn.py:124-142)
    -> `_, idx = torch.sort(outputs, descending=True); idx[:10]` (Helpers/Metrics.py:60-61).
    With `candidates` [B, C] only those items are scored (BASELINE.json configs[4]) and the
    returned ids are item ids.  Returns (ids int64 [B, k], scores [B, k])."""
    with torch.no_grad():
        f = model.features() if features is None else features
        ids, vals = [], []
        for b in range(int(queries.numel())):
            if candidates is None:
                ones = torch.ones(model.I, dtype=torch.long)
                out = model.forward(users[b] * ones, queries[b] * ones, None, features=f)      # all items
                _, order = torch.sort(out, descending=True, stable=True)
                top = order[:k]
                ids.append(top)
                vals.append(out[top])
            else:
                c = candidates[b]
                ones = torch.ones(c.numel(), dtype=torch.long)
                out = model.forward(users[b] * ones, queries[b] * ones, c, features=f)
                _, order = torch.sort(out, descending=True, stable=True)
                top = order[:k]
                ids.append(c[top])
                vals.append(out[top])
        return torch.stack(ids), torch.stack(vals)


def metrics_at_10(recommend_indices: Sequence[int], interacted_items: Sequence[int]):
    """Helpers/Metrics.py:47-88 for flags_are_all_1 (the only form TestSearchLogDataLoader emits,
    Dataset.py:312): (HR@10, NDCG@10, MAP@10) of one search from its top-10 item ids."""
    import math
    rec = [int(x) for x in recommend_indices][:10]
    hits = [rec.index(int(it)) for it in interacted_items if int(it) in rec]        # :66-68
    n10 = min(len(interacted_items), 10)                                             # :62
    hr = len(hits) / n10                                                             # :80
    dcg = sum(math.log(2, i + 2) for i in hits)                                      # :93
    idcg = sum(math.log(2, i + 2) for i in range(n10))                               # :97-103
    ap = sum((j + 1) / (i + 1) for j, i in enumerate(hits)) / len(hits) if hits else 0.0   # :105-109 (hit order = interacted_items order, as the reference)
    return hr, dcg / idcg, ap


def bce_with_logits_mean(scores: torch.Tensor, flags: torch.Tensor) -> torch.Tensor:
    """Main.py:191 / TrainTestHelper.py:132: nn.BCEWithLogitsLoss() (mean)."""
    return torch.nn.functional.binary_cross_entropy_with_logits(scores, flags)


def conv_fwd_bwd(model: OracleModel, x: torch.Tensor) -> torch.Tensor:
    """Metric M1's unit of work: fwd+bwd through the L-layer stack with
    loss = sum(cat(outs, 1)); returns dX.  Used by bench.py's CPU baseline legs."""
    x = x.detach().clone().requires_grad_(True)
    for p in model.params.values():
        p.grad = None
    outs = model.conv_stack(x)
    torch.cat(outs, 1).sum().backward()
    return x.grad
