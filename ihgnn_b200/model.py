"""`RawGnn`: the caller of the hot path, mirrored from /root/reference/Models/RawGnn.py so the
stack can run where the reference tree is absent (the GPU box, bench.py, smoke()).

Same constructor keywords, module names (`embeddings`, `gnn_{k}`, `prediction_layer`) and
hence `state_dict` keys as RawGnn.py:14-101; `forward` / `save_features_for_test` follow
:104-155.  The reference's own RawGnn also runs unchanged on top of these layers
(INTEGRATION.md); the only liberties taken here are the ones the drop-in layers permit: the
three embedding blocks come from one fused kernel instead of `torch.cat`, and the batch row
selects use the library's gather kernel (deterministic backward).
"""
from __future__ import annotations

from typing import Optional, Type

import torch
import torch.nn as nn
from torch import Tensor

from . import functional as F_
from .layers import EmbeddingLayer, GCNLayer, HGCNLayer, HemPredictionLayer, IHGNNLayer
from .settings import Gs


class RawGnn(nn.Module):

    _saved_output_feature: Optional[Tensor] = None

    def __init__(self, device, dataset, embedding_size: int, gnn_layer_type: Type,
                 gnn_layer_count: int, feature_interaction_order: int, phase2_attention: bool,
                 predictions: Type, lambda_muq: float):
        super().__init__()
        self.device = device
        self.dataset = dataset
        self.embedding_size = embedding_size
        self.gnn_layer_type = gnn_layer_type
        self.gnn_layer_count = gnn_layer_count
        self.feature_interaction_order = feature_interaction_order
        self.phase2_attention = phase2_attention
        self.prediction_layer_type = predictions
        self.output_feature_size = embedding_size * (1 + gnn_layer_count)

        self.embeddings = EmbeddingLayer(dataset=dataset, embedding_size=embedding_size)
        self.gnns = []
        for layer in range(gnn_layer_count):
            if gnn_layer_type in (HGCNLayer, GCNLayer):                        # RawGnn.py:61-73
                self.gnns.append(gnn_layer_type(device=device, dataset=dataset, input_dimension=embedding_size,
                                                output_dimension=embedding_size))
            elif gnn_layer_type is IHGNNLayer:
                order = feature_interaction_order
                if order > 1 and layer > 0:                  # RawGnn.py:76-78
                    order = 1
                self.gnns.append(IHGNNLayer(device=device, dataset=dataset, input_dimension=embedding_size,
                                            output_dimension=embedding_size,
                                            feature_interaction_order=order,
                                            phase2_attention=phase2_attention))
            else:
                raise NotImplementedError(f"unsupported GNN layer type: {gnn_layer_type}")
        for i, gnn in enumerate(self.gnns):
            self.add_module(f"gnn_{i}", gnn)
        if predictions is HemPredictionLayer:
            self.prediction_layer = HemPredictionLayer(feature_dimension=self.output_feature_size,
                                                       lambda_muq=lambda_muq, item_count=dataset.item_count)
        else:
            raise NotImplementedError(f"unsupported prediction layer type: {predictions}")

    # ---- conv stack -------------------------------------------------------------------
    def conv_stack(self, input_features: Tensor):
        """RawGnn.py:113-118: list of the input and every layer's output."""
        outs = [input_features]
        h = input_features
        for gnn in self.gnns:
            h = gnn(h)
            outs.append(h)
        return outs

    def output_features(self) -> Tensor:
        """F = cat(all layer outputs, 1): [N, d(1+L)]  (RawGnn.py:110-122)."""
        return torch.cat(self.conv_stack(self.embeddings.embed_all()), 1)

    def forward(self, user_indices: Tensor, query_indices: Tensor, item_indices: Optional[Tensor] = None):
        ds = self.dataset
        if self._saved_output_feature is None:
            # training: cat(gnn_outputs, 1)[rows] == cat([o[rows] for o in gnn_outputs], 1)  (RawGnn.py:121-133):
            # gather the batch rows out of every layer's table and concatenate the small results -- the
            # [N, d(1+L)] table and its dense gradient are never materialised
            idxs = [user_indices, query_indices] + ([item_indices] if item_indices is not None else [])
            offs = [0, ds.query_start_index_in_graph, ds.item_start_index_in_graph][:len(idxs)]
            # every layer output is tapped for the batch rows and passed on to the next layer; the tap's backward
            # adds the batch-row gradients into the dense gradient coming back from that layer (F_.TapRowsFn)
            h, rows = F_.tap_rows(self.embeddings.embed_all(), idxs, offs)
            outs, picked = [h], [rows]
            for gnn in self.gnns:
                h, rows = F_.tap_rows(gnn(h), idxs, offs)
                outs.append(h)
                picked.append(rows)
            fu = torch.cat([p[0] for p in picked], 1)                                      # :128
            fq = torch.cat([p[1] for p in picked], 1)                                      # :129
            if item_indices is not None:
                fi = torch.cat([p[2] for p in picked], 1)                                  # :131
            else:
                fi = torch.cat([o[ds.item_start_index_in_graph:] for o in outs], 1)        # :133
            return self.prediction_layer(fu, fq, fi, item_indices)
        output_feature = self._saved_output_feature
        fu = F_.gather_rows(output_feature, user_indices, 0)
        fq = F_.gather_rows(output_feature, query_indices, ds.query_start_index_in_graph)
        if item_indices is not None:
            fi = F_.gather_rows(output_feature, item_indices, ds.item_start_index_in_graph)
        else:
            fi = output_feature[ds.item_start_index_in_graph:]
        return self.prediction_layer(fu, fq, fi, item_indices)

    @torch.no_grad()
    def rank(self, user_indices: Tensor, query_indices: Tensor, candidates: Optional[Tensor] = None,
             k: int = 10):
        """Batched evaluation: the k best items per (user, query) -- all items, or the given
        candidate lists [B, C] -- in one launch on the saved features.  Equals, per query b,
        `torch.sort(self(u_b * ones(I), q_b * ones(I), None), descending=True)[1][:k]`
        (TrainTestHelper.py:56 + Metrics.py:60-61).  Returns (item ids [B, k], scores [B, k])."""
        ds = self.dataset
        feat = self._saved_output_feature
        if feat is None:
            feat = self.output_features()
        p = self.prediction_layer
        return F_.rank_topk(feat, user_indices, query_indices, p.items_bias, p.lambda_muq,
                            query_row0=ds.query_start_index_in_graph, item_row0=ds.item_start_index_in_graph,
                            item_count=ds.item_count, candidates=candidates, k=k,
                            cosine=bool(Gs.Prediction.use_cosine_similarity))

    def save_features_for_test(self) -> None:
        self._saved_output_feature = self.output_features()

    def clear_saved_feature(self) -> None:
        self._saved_output_feature = None


def rank_searches(model, user_indices: Tensor, query_indices: Tensor, candidates: Optional[Tensor] = None,
                  k: int = 10):
    """`RawGnn.rank` for any RawGnn-shaped model -- this package's or the reference's own
    (Models/RawGnn.py) running on the drop-in layers: needs `_saved_output_feature`
    (`save_features_for_test()` must have run), `dataset` offsets and a HemPredictionLayer."""
    feat = model._saved_output_feature
    if feat is None:
        raise RuntimeError("rank_searches: call model.save_features_for_test() first")
    ds, p = model.dataset, model.prediction_layer
    return F_.rank_topk(feat, user_indices, query_indices, p.items_bias, p.lambda_muq,
                        query_row0=ds.query_start_index_in_graph, item_row0=ds.item_start_index_in_graph,
                        item_count=ds.item_count, candidates=candidates, k=k,
                        cosine=bool(Gs.Prediction.use_cosine_similarity))


def metrics_at_10(recommended, interacted_items):
    """(HR@10, NDCG@10, MAP@10) of one search from its top-10 item ids: the arithmetic of
    `Metrics.calculate_on_all_items` for flags_are_all_1 (Helpers/Metrics.py:60-110), including its
    MAP convention (hits enumerated in `interacted_items` order, :105-109).  Pure host code."""
    import math
    rec = [int(x) for x in recommended][:10]
    items = [int(x) for x in interacted_items]
    hits = [rec.index(it) for it in items if it in rec]                                 # Metrics.py:66-68
    n10 = min(len(items), 10)                                                           # :62
    hr = len(hits) / n10                                                                # :80
    ndcg = sum(math.log(2, i + 2) for i in hits) / sum(math.log(2, i + 2) for i in range(n10))   # :84, :93-103
    ap = sum((j + 1) / (i + 1) for j, i in enumerate(hits)) / len(hits) if hits else 0.0          # :81, :105-109
    return hr, ndcg, ap


def search_metrics(model, logs, batch_size: int = 8192):
    """Per-search (HR@10, NDCG@10, MAP@10) of `logs` = sequence of (user, query, interacted_items, ...)
    tuples (TestSearchLogDataLoader.logs, Dataset.py:297-318), every search ranked against all items
    in batches of `batch_size` searches per launch; the metric arithmetic is Helpers/Metrics.py:47-110
    for flags_are_all_1 (the only form the reference's loader emits).  `model` must hold saved features."""
    dev = model._saved_output_feature.device
    out = []
    for s in range(0, len(logs), batch_size):
        part = logs[s:s + batch_size]
        users = torch.tensor([l[0] for l in part], dtype=torch.int64).to(dev)
        queries = torch.tensor([l[1] for l in part], dtype=torch.int64).to(dev)
        top, _ = rank_searches(model, users, queries, None, 10)
        top = top.cpu().tolist()                               # one D2H copy per batch
        out.extend(metrics_at_10(rec, l[2]) for rec, l in zip(top, part))
    return out


def evaluate_searches(model, logs, batch_size: int = 8192, k: int = 10):
    """Batched form of `test_and_get_avg_metrics` (Helpers/TrainTestHelper.py:37-102) for RawGnn
    models: features are saved once, every search is ranked against all items (one launch per
    `batch_size` searches instead of one forward + full sort + .cpu() per search) and HR@10 /
    NDCG@10 / MAP@10 are averaged over the searches.  Returns (hit_ratio, ndcg, map)."""
    logs = [l for l in logs if len(l[2]) > 0]
    if not logs:
        return 0.0, 0.0, 0.0
    with torch.no_grad():
        model.save_features_for_test()
        try:
            per = search_metrics(model, logs, batch_size)
        finally:
            model.clear_saved_feature()
    n = len(per)
    return sum(p[0] for p in per) / n, sum(p[1] for p in per) / n, sum(p[2] for p in per) / n


def make_fast_test_and_get_avg_metrics(original, metrics_cls, log_print=None):
    """A drop-in for the reference's `test_and_get_avg_metrics(model, dataset_train, dataloader,
    get_long_tail_stat=False)` (Helpers/TrainTestHelper.py:37-102) with the same return value --
    (per-user averages or None, average Metrics, seconds) -- that ranks the searches in batches on the
    GPU.  `metrics_cls` is the reference's `Helpers.Metrics.Metrics`; models that are not RawGnn-shaped
    (Srrl: no `_saved_output_feature`) and logs whose flags are not all 1 go to `original`."""
    import time

    def test_and_get_avg_metrics(model, dataset_train, dataloader, get_long_tail_stat: bool = False):
        # RawGnn-shaped models only: Srrl also has save_features_for_test / prediction_layer but saves
        # `_saved_u_ps` etc. (Models/Srrl.py), and logs with flags above 1 need the graded NDCG of
        # Metrics.py:84-103 -- both go to the reference's own loop
        if not (hasattr(model, "save_features_for_test") and hasattr(model, "prediction_layer")
                and hasattr(model, "_saved_output_feature") and hasattr(dataloader, "logs")
                and all(len(l) < 5 or bool(l[4]) for l in dataloader.logs)):
            return original(model, dataset_train, dataloader, get_long_tail_stat)
        start = time.time()
        logs = dataloader.logs
        total = metrics_cls()
        per_user = [[] for _ in range(dataset_train.user_count)] if get_long_tail_stat else None
        with torch.no_grad():
            model.save_features_for_test()                                              # :50-51
            try:
                per = search_metrics(model, logs)
            finally:
                model.clear_saved_feature()                                             # :87-88
        for l, (hr, ndcg, ap) in zip(logs, per):
            m = metrics_cls()
            m.HitRatio_at10, m.NDCG_at10, m.MAP_at10 = hr, ndcg, ap
            total.add_to_self(m)                                                        # :65-66
            if per_user is not None:
                per_user[int(l[0])].append(m)                                           # :69-70
        u_metrics = None
        if per_user is not None:                                                        # :72-81
            u_metrics = []
            for ms in per_user:
                if not ms:
                    u_metrics.append(None)
                else:
                    m0 = metrics_cls()
                    for m in ms:
                        m0.add_to_self(m)
                    u_metrics.append(m0.divide_and_get_new(len(ms)))
        avg = total.divide_and_get_new(len(logs))                                       # :91
        elapsed = time.time() - start
        if log_print is not None:
            log_print(f"test done in {elapsed:<.2f} s, {len(logs)} search logs (batched GPU ranking).")
            log_print(avg.to_string(highlight=True))
        return u_metrics, avg, elapsed

    return test_and_get_avg_metrics
