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
from .layers import EmbeddingLayer, HGCNLayer, HemPredictionLayer, IHGNNLayer


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
            if gnn_layer_type is HGCNLayer:
                self.gnns.append(HGCNLayer(device=device, dataset=dataset, input_dimension=embedding_size,
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
            outs = self.conv_stack(self.embeddings.embed_all())
            idxs = [user_indices, query_indices] + ([item_indices] if item_indices is not None else [])
            offs = [0, ds.query_start_index_in_graph, ds.item_start_index_in_graph][:len(idxs)]
            picked = [F_.gather_rows_multi(o, idxs, offs) for o in outs]           # one dense gradient per table
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

    def save_features_for_test(self) -> None:
        self._saved_output_feature = self.output_features()

    def clear_saved_feature(self) -> None:
        self._saved_output_feature = None
