"""Drop-in layer classes: same constructor / forward signatures, submodule names and
`state_dict` keys as the reference, computed by the sm_100a kernels of
libihgnn_b200.so.  CUDA only -- a CPU tensor raises.

  EmbeddingLayer       /root/reference/Models/EmbeddingLayers.py:11-104
  FeatureInteractor    /root/reference/Models/CommonLayers.py:29-87
  IHGNNLayer           /root/reference/Models/GnnLayers.py:156-236
  HGCNLayer            /root/reference/Models/GnnLayers.py:118-153
  HemPredictionLayer   /root/reference/Models/PredictionLayers.py:6-44
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.init as init
from torch import Tensor
from torch.nn.parameter import Parameter

from . import _lib
from . import functional as F_
from .graph import CsrPlan, Pps2DGraph, PpsHyperGraph, csr_from_keys
from .settings import Gs, Gsv


# --------------------------------------------------------------------------------------
# embeddings
# --------------------------------------------------------------------------------------
class _BagTables:
    """Static index structures of the query EmbeddingBag: the bag CSR (query -> word ids) and
    its stable transpose (word -> queries), both with segment plans."""

    def __init__(self, dataset, vocab_rows: int):
        words = dataset.queries_for_embeddingbag            # int64, already +1  (Dataset.py:169)
        offsets = dataset.queries_offset_for_embeddingbag   # int64 [Q] start offsets
        _lib.require_cuda(words, offsets)
        dev = words.device
        self.user_count = int(len(dataset.users_onehot))
        self.item_count = int(len(dataset.items_onehot))
        Q = int(offsets.numel())
        self.query_count = Q
        nnz = int(words.numel())
        ptr = torch.empty(Q + 1, dtype=torch.int32, device=dev)
        ptr[:Q] = offsets.to(torch.int32)
        ptr[Q] = nnz
        lens = (ptr[1:] - ptr[:-1]).to(torch.float32)
        # EmbeddingBag(mean): sum / len, empty bag -> zeros
        self.bag_inv_len = torch.where(lens > 0, 1.0 / lens.clamp(min=1.0), torch.zeros_like(lens))
        words32 = words.to(torch.int32).contiguous()
        self.bag_plan = CsrPlan(ptr, words32)
        # transpose: for every word the (stably ordered) list of bags it occurs in
        bag_of_pos = torch.repeat_interleave(torch.arange(Q, device=dev, dtype=torch.int32),
                                             (ptr[1:] - ptr[:-1]).to(torch.int64))
        wptr, _perm, bags = csr_from_keys(words32, vocab_rows, values=bag_of_pos)
        self.word_plan = CsrPlan(wptr, bags)


class EmbeddingLayer(nn.Module):
    """User / item embedding tables (+1 padding row) and the query EmbeddingBag(mean)."""

    def __init__(self, dataset, embedding_size: int):
        super().__init__()
        self.dataset = dataset
        self.embedding_size = embedding_size
        self.users = dataset.users_onehot
        self.queries = dataset.queries_multihot
        self.queries_bag = dataset.queries_for_embeddingbag
        self.queries_bag_offset = dataset.queries_offset_for_embeddingbag
        self.items = dataset.items_onehot
        self.vocabulary = dataset.vocabulary_onehot
        # same construction (hence RNG draw) order as EmbeddingLayers.py:33-35
        self.embedding_user = EmbeddingLayer.create_embedding(len(self.users) + 1, embedding_size, padding_idx=0)
        self.embedding_item = EmbeddingLayer.create_embedding(len(self.items) + 1, embedding_size, padding_idx=0)
        self.embedding_bag_vocabulary = EmbeddingLayer.create_embedding_bag(len(self.vocabulary) + 1, embedding_size)
        if Gs.Query.transform == Gsv.mean:                  # EmbeddingLayers.py:38-48
            pass
        elif Gs.Query.transform == Gsv.activation:
            self.query_transform = nn.Sequential(nn.Linear(embedding_size, embedding_size),
                                                 Gs.Query.transform_activation())
        elif Gs.Query.transform == Gsv.rnn:
            raise NotImplementedError()                     # as the reference (:45-46)
        else:
            raise ValueError()
        self._tables: Optional[_BagTables] = None

    @property
    def tables(self) -> _BagTables:
        if self._tables is None:
            self._tables = _BagTables(self.dataset, len(self.vocabulary) + 1)
        return self._tables

    def forward(self, user_indices: Tensor = None, query_indices: Tensor = None, item_indices: Tensor = None):
        if user_indices is None and query_indices is None and item_indices is None \
                and self.embedding_bag_vocabulary is not None:
            x = self.embed_all()
            U, Q = self.tables.user_count, self.tables.query_count
            return x[:U], x[U:U + Q], x[U + Q:]
        return self.embed_user(user_indices), self.embed_query(query_indices), self.embed_item(item_indices)

    def embed_all(self) -> Tensor:
        """[users; queries; items] feature matrix [N, d] in one pass (what RawGnn.py:112 cats)."""
        if Gs.Query.transform == Gsv.activation:            # the transformed query rows are a separate tensor
            return torch.cat([self.embed_user(), self.embed_query(), self.embed_item()])
        return F_.EmbedAllFn.apply(self.embedding_user.weight, self.embedding_bag_vocabulary.weight,
                                   self.embedding_item.weight, self.tables)

    def embed_user(self, user_indices: Tensor = None) -> Tensor:
        w = self.embedding_user.weight
        _lib.require_cuda(w)
        if user_indices is None:
            return w[1:]                                    # identity index 1..U: a view
        return F_.gather_rows(w, user_indices, 1)           # one-hot == index + 1

    def embed_item(self, item_indices: Tensor = None) -> Tensor:
        w = self.embedding_item.weight
        _lib.require_cuda(w)
        if item_indices is None:
            return w[1:]
        return F_.gather_rows(w, item_indices, 1)

    def embed_query(self, query_indices: Tensor = None) -> Tensor:
        t = self.tables
        w = self.embedding_bag_vocabulary.weight
        _lib.require_cuda(w)
        q = _BagMeanFn.apply(w, t)
        if query_indices is not None:                       # EmbeddingLayers.py:80-81
            q = F_.gather_rows(q, query_indices, 0)
        if Gs.Query.transform == Gsv.activation:            # :83-84: Linear on the node-Linear kernel, then the activation module
            lin, act = self.query_transform[0], self.query_transform[1]
            q = act(F_.typed_linear(q, lin.weight.unsqueeze(0), lin.bias.unsqueeze(0), None))
        return q

    @staticmethod
    def create_embedding(num_embeddings: int, embedding_dimension: int, padding_idx: int = None) -> nn.Embedding:
        emb = nn.Embedding(num_embeddings, embedding_dimension, padding_idx=padding_idx)
        init.xavier_uniform_(emb.weight)
        return emb

    @staticmethod
    def create_embedding_bag(num_embeddings: int, embedding_dimension: int, mode: str = "mean") -> nn.EmbeddingBag:
        emb = nn.EmbeddingBag(num_embeddings, embedding_dimension, mode=mode)
        init.xavier_uniform_(emb.weight)
        return emb


class _BagMeanFn(torch.autograd.Function):
    """EmbeddingBag(mean) over all Q queries (EmbeddingLayers.py:79)."""

    @staticmethod
    def forward(ctx, w_vocab, tables: _BagTables):
        ctx.tables = tables
        w_vocab = _lib.rows_f32(w_vocab)
        return F_.segment_reduce(tables.bag_plan, w_vocab, int(w_vocab.shape[1]), row_scale=tables.bag_inv_len)

    @staticmethod
    def backward(ctx, dq):
        t = ctx.tables
        dq = _lib.rows_f32(dq)
        return F_.segment_reduce(t.word_plan, dq, int(dq.shape[1]), src_scale=t.bag_inv_len), None


# --------------------------------------------------------------------------------------
# node -> hyperedge -> node primitives with autograd
# --------------------------------------------------------------------------------------
class _EdgeGatherSumFn(torch.autograd.Function):
    """ef[e] = alpha * sum_{n in e} node_scale[n] * h[n];  backward is the CSR segmented sum."""

    @staticmethod
    def forward(ctx, h, graph: PpsHyperGraph, node_scale, alpha: float, bwd_row_scale):
        ctx.graph, ctx.bwd_row_scale = graph, bwd_row_scale
        return F_.edge_gather_sum(h, graph.i3, node_scale=node_scale, alpha=alpha)

    @staticmethod
    def backward(ctx, def_):
        g = ctx.graph
        def_ = _lib.rows_f32(def_)
        dh = F_.segment_reduce(g.plan, def_, int(def_.shape[1]), row_scale=ctx.bwd_row_scale)
        return dh, None, None, None, None


class _TwoHopFn(torch.autograd.Function):
    """out[r] = row_scale[r] * alpha * sum_{e contains r} sum_{n in e} node_scale[n] * h[n]:
    _EdgeGatherSumFn followed by _ScatterMeanFn in one pass over the node table (no [E,d]
    intermediate).  H H^T is symmetric, so backward is the same kernel with the scales swapped."""

    @staticmethod
    def forward(ctx, h, graph: PpsHyperGraph, node_scale, alpha: float, row_scale, own=(1.0, 0.0)):
        ctx.graph, ctx.node_scale, ctx.alpha, ctx.row_scale, ctx.own = graph, node_scale, alpha, row_scale, own
        return F_.two_hop_reduce(graph.plan, _two_hop_nbr(graph), h, node_scale=node_scale, alpha=alpha,
                                 row_scale=row_scale, own=own)

    @staticmethod
    def backward(ctx, dout):
        g = ctx.graph
        dh = F_.two_hop_reduce(g.plan, _two_hop_nbr(g), dout, node_scale=ctx.row_scale, alpha=ctx.alpha,
                               row_scale=ctx.node_scale, own=ctx.own)
        return dh, None, None, None, None, None


class _PairAdjFn(torch.autograd.Function):
    """out = D^-1/2 (A [+ I]) D^-1/2 h over the explicit pair CSR of `Pps2DGraph` (graph_only_* completeness,
    integer flag weights: Helpers/Graph.py:40-65); A is symmetric, so backward is the same product."""

    @staticmethod
    def _apply(h, g2):
        h = _lib.rows_f32(h)
        s = g2.dv_inv_sqrt
        init = h * s.view(-1, 1) if g2.use_self_connection else None
        return F_.segment_reduce(g2.pair_plan, h, int(h.shape[1]), src_scale=s, row_scale=s, init=init)

    @staticmethod
    def forward(ctx, h, g2):
        ctx.g2 = g2
        return _PairAdjFn._apply(h, g2)

    @staticmethod
    def backward(ctx, dout):
        return _PairAdjFn._apply(dout, ctx.g2), None


def _two_hop_nbr(graph) -> Tensor:
    return graph.plan.two_hop_nbr(graph.i3, graph.type_bounds, getattr(graph, "row_slot", None))


def _gather_scatter(h: Tensor, graph, node_scale, alpha: float, gather_bwd_scale, row_scale) -> Tensor:
    """row_scale * H . (alpha * H^T . (node_scale * h)): one two-hop pass when the node table is
    L2-resident, else gather-sum into [E,d] + segmented reduce."""
    if F_.two_hop_enabled(int(h.shape[0]), int(h.shape[1])):
        return _TwoHopFn.apply(h, graph, node_scale, alpha, row_scale)
    ef = _EdgeGatherSumFn.apply(h, graph, node_scale, alpha, gather_bwd_scale)
    return _ScatterMeanFn.apply(ef, graph, row_scale)


class _EdgeInteractFn(torch.autograd.Function):
    """Order 2/3 hyperedge features (CommonLayers.py:68-85) with the first-order blocks hoisted:
    ef[e] = p[u]+p[q]+p[i] + W_hi . cat(u*q, q*i, i*u [, u*q*i])."""

    @staticmethod
    def forward(ctx, xp, p, w_hi, graph: PpsHyperGraph, order: int):
        xp, p, w_hi = _lib.rows_f32(xp), _lib.rows_f32(p), _lib.rows_f32(w_hi)
        dim, E = int(xp.shape[1]), graph.EdgeCount
        ef = torch.empty((E, dim), dtype=torch.float32, device=xp.device)
        ws_bytes = _lib.lib().ihg_edge_interact_fwd_workspace_bytes(dim, order)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=xp.device)
        _lib.call("ihg_edge_interact_fwd", _lib.ptr(xp), _lib.ld(xp), _lib.ptr(p), _lib.ld(p),
                  _lib.ptr(w_hi), _lib.ld(w_hi), order, _lib.ptr(graph.i3), E, _lib.ptr(ef), dim,
                  dim, _lib.ptr(ws), ws_bytes, _lib.stream_ptr(),
                  tag="edge_interact_fwd", algo_bytes=E * (12 + 28 * dim))
        ctx.graph, ctx.order = graph, order
        ctx.save_for_backward(xp, w_hi)
        return ef

    @staticmethod
    def backward(ctx, def_):
        xp, w_hi = ctx.saved_tensors
        g, order = ctx.graph, ctx.order
        def_ = _lib.rows_f32(def_)
        dim, E = int(xp.shape[1]), g.EdgeCount
        nb = 4 if order == 3 else 3
        dp = F_.segment_reduce(g.plan, def_, dim)
        slot_grad = torch.empty((E, 3, dim), dtype=torch.float32, device=xp.device)
        dw_hi = torch.empty((dim, nb * dim), dtype=torch.float32, device=xp.device)
        ws_bytes = _lib.lib().ihg_edge_interact_bwd_workspace_bytes(dim, order)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xp.device)
        _lib.call("ihg_edge_interact_bwd", _lib.ptr(xp), _lib.ld(xp), _lib.ptr(def_), _lib.ld(def_),
                  _lib.ptr(w_hi), _lib.ld(w_hi), order, _lib.ptr(g.i3), E, _lib.ptr(slot_grad),
                  _lib.ptr(dw_hi), dim, _lib.ptr(ws), ws_bytes, _lib.stream_ptr(),
                  tag="edge_interact_bwd", algo_bytes=E * (12 + 28 * dim))
        dxp = F_.segment_reduce(g.plan, slot_grad, dim, src_row_mul=3, bounds=g.type_bounds,
                                row_slot=getattr(g, "row_slot", None))
        return dxp, dp, dw_hi, None, None


class _FeatureInteractFn(torch.autograd.Function):
    """The whole order-2/3 FeatureInteractor.forward (CommonLayers.py:68-85) in one tensor-core
    kernel: ef = aggregation.weight . cat(u, q, i, u*q, q*i, i*u [, u*q*i]) + bias, the raw rows
    being three more operand blocks (no first-order table, no gather in the epilogue).
    Backward uses the hoisted algebra: dP = H^T-reduce(def) per node, then the typed node Linear
    backward for the first-order blocks, plus the product-rule kernels for the rest."""

    @staticmethod
    def forward(ctx, xp, w_agg, bias, graph: PpsHyperGraph, order: int):
        xp, w_agg = _lib.rows_f32(xp), _lib.rows_f32(w_agg)
        bias = bias.contiguous()
        dim, E = int(xp.shape[1]), graph.EdgeCount
        ef = torch.empty((E, dim), dtype=torch.float32, device=xp.device)
        ws_bytes = _lib.lib().ihg_edge_interact_fwd_workspace_bytes(dim, order)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=xp.device)
        _lib.call("ihg_feature_interact_fwd", _lib.ptr(xp), _lib.ld(xp), _lib.ptr(w_agg), _lib.ld(w_agg),
                  _lib.ptr(bias), order, _lib.ptr(graph.i3), E, _lib.ptr(ef), dim, dim, _lib.ptr(ws),
                  ws_bytes, _lib.stream_ptr(), tag="edge_interact_fwd", algo_bytes=E * (12 + 16 * dim))
        ctx.graph, ctx.order = graph, order
        ctx.save_for_backward(xp, w_agg)
        return ef

    @staticmethod
    def backward(ctx, def_):
        xp, w_agg = ctx.saved_tensors
        g, order = ctx.graph, ctx.order
        def_ = _lib.rows_f32(def_)
        dim, E = int(xp.shape[1]), g.EdgeCount
        nb = 4 if order == 3 else 3
        w_hi = w_agg[:, 3 * dim:]
        w_lo = _split_first_order(w_agg, dim).contiguous()              # [3, dim, dim]
        dp = F_.segment_reduce(g.plan, def_, dim)
        slot_grad = torch.empty((E, 3, dim), dtype=torch.float32, device=xp.device)
        dw_hi = torch.empty((dim, nb * dim), dtype=torch.float32, device=xp.device)
        ws_bytes = _lib.lib().ihg_edge_interact_bwd_workspace_bytes(dim, order)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xp.device)
        _lib.call("ihg_edge_interact_bwd", _lib.ptr(xp), _lib.ld(xp), _lib.ptr(def_), _lib.ld(def_),
                  _lib.ptr(w_hi), _lib.ld(w_hi), order, _lib.ptr(g.i3), E, _lib.ptr(slot_grad),
                  _lib.ptr(dw_hi), dim, _lib.ptr(ws), ws_bytes, _lib.stream_ptr(),
                  tag="edge_interact_bwd", algo_bytes=E * (12 + 28 * dim))
        dxp_hi = F_.segment_reduce(g.plan, slot_grad, dim, src_row_mul=3, bounds=g.type_bounds)
        # first-order blocks: dX' = dP . W_a[:, slot] (typed by node type) + the product-rule part
        dxp = F_.node_linear(dp, w_lo, transpose_w=True, addend=dxp_hi, bounds=g.type_bounds)
        dw_lo, db_lo = F_.node_linear_wgrad(dp, xp, 3, g.type_bounds, True)
        dw = torch.cat([dw_lo[0], dw_lo[1], dw_lo[2], dw_hi], 1)
        # every hyperedge has exactly one user: sum_e def[e] == sum over user rows of dP
        return dxp, dw, db_lo[0], None, None


class _ScatterMeanFn(torch.autograd.Function):
    """out[v] = row_scale[v] * sum_{e contains v} ef[e]   (thsp.matmul(incidence, ef) * Dv^-1,
    GnnLayers.py:233-234); backward is the node -> hyperedge gather-sum."""

    @staticmethod
    def forward(ctx, ef, graph: PpsHyperGraph, row_scale):
        ctx.graph, ctx.row_scale = graph, row_scale
        ef = _lib.rows_f32(ef)
        return F_.segment_reduce(graph.plan, ef, int(ef.shape[1]), row_scale=row_scale)

    @staticmethod
    def backward(ctx, dout):
        g = ctx.graph
        return F_.edge_gather_sum(dout, g.i3, node_scale=ctx.row_scale), None, None


def _split_first_order(weight: Tensor, d: int) -> Tensor:
    """aggregation.weight[:, :3d] ([d_out, 3d]) -> per-slot stack [3, d_out, d]."""
    return weight[:, :3 * d].reshape(weight.shape[0], 3, d).permute(1, 0, 2)


# --------------------------------------------------------------------------------------
# FeatureInteractor / IHGNN / HGCN
# --------------------------------------------------------------------------------------
class FeatureInteractor(nn.Module):
    """node features [N,d] -> hyperedge features [E,d_out] (CommonLayers.py:29-87)."""

    def __init__(self, dataset, max_order: int, node_feature_dimension: int, output_dimension: int):
        super().__init__()
        self.max_order = max_order
        self.node_feature_dimension = node_feature_dimension
        self.output_dimension = output_dimension
        self.graph: PpsHyperGraph = dataset.graph
        if max_order == 1:
            self.aggregation = nn.Linear(3 * node_feature_dimension, output_dimension)
        elif max_order in (2, 3):
            self.aggregation = nn.Linear((6 if max_order == 2 else 7) * node_feature_dimension, output_dimension)
        else:
            raise ValueError(f"max_order must be 1, 2 or 3 (got {max_order})")

    def _first_order(self, node_features: Tensor) -> Tensor:
        """p = X' W_a[:, slot]^T per node type, aggregation bias folded into the user slot."""
        d = self.node_feature_dimension
        w_lo = _split_first_order(self.aggregation.weight, d)
        zeros = torch.zeros_like(self.aggregation.bias)
        b_lo = torch.stack([self.aggregation.bias, zeros, zeros])
        return F_.typed_linear(node_features, w_lo, b_lo, self.graph.type_bounds)

    def forward(self, node_features: Tensor) -> Tensor:
        _lib.require_cuda(node_features)
        g, d = self.graph, self.node_feature_dimension
        if self._full_supported():
            return _FeatureInteractFn.apply(node_features, self.aggregation.weight, self.aggregation.bias,
                                            g, self.max_order)
        p = self._first_order(node_features)
        if self.max_order == 1:
            return _EdgeGatherSumFn.apply(p, g, None, 1.0, None)
        if self.output_dimension != d:
            raise NotImplementedError("order 2/3 FeatureInteractor requires output_dimension == node_feature_dimension")
        return _EdgeInteractFn.apply(node_features, p, self.aggregation.weight[:, 3 * d:], g, self.max_order)

    def _full_supported(self) -> bool:
        return (self.max_order in (2, 3) and self.output_dimension == self.node_feature_dimension
                and bool(_lib.lib().ihg_feature_interact_supported(self.node_feature_dimension)))


class IHGNNLayer(nn.Module):
    """out = Dv^-1 . H . FeatureInteractor(Linear(X))      (GnnLayers.py:221-236)."""

    def __init__(self, device, dataset, input_dimension: int, output_dimension: int,
                 feature_interaction_order: int, phase2_attention: bool):
        super().__init__()
        self.device = device
        self.dataset = dataset
        self.feature_interaction_order = feature_interaction_order
        self.attention_phase2 = phase2_attention
        assert feature_interaction_order in [1, 2, 3], "feature_interaction_order must be 1, 2 or 3"
        if phase2_attention:
            # dead and broken in the reference (Main.py:57; GnnLayers.py:161-164 vs :62,90)
            raise NotImplementedError("phase2_attention is not supported (dead code path in the reference)")
        graph: PpsHyperGraph = dataset.graph          # reference: dataset.hypergraph (the same object there; GnnLayers.py:185)
        self.graph = graph
        self.Dv_neg_1 = graph.dv_inv.view(-1, 1)            # GnnLayers.py:187
        # aggregation Linear first, then the projection Linear: same RNG order as :193,:218
        self.feature_interactor = FeatureInteractor(dataset=dataset, max_order=feature_interaction_order,
                                                    node_feature_dimension=input_dimension,
                                                    output_dimension=input_dimension)
        self.feature_transform = nn.Linear(input_dimension, output_dimension)

    def forward(self, input_features: Tensor) -> Tensor:
        _lib.require_cuda(input_features)
        g = self.graph
        fi = self.feature_interactor
        wt, bt = self.feature_transform.weight, self.feature_transform.bias
        if self.feature_interaction_order == 1:
            # Fold the two Linears at node level: W_a[:,s] (W_t x + b_t) = (W_a[:,s] W_t) x + W_a[:,s] b_t
            # (two d x d products on the parameters instead of a second [N,d]x[d,d] pass)
            d = fi.node_feature_dimension
            w_lo = _split_first_order(fi.aggregation.weight, d)                 # [3, d, d]
            w_f = torch.matmul(w_lo, wt)                                         # [3, d, d_in]
            b_f = torch.matmul(w_lo, bt)                                         # [3, d]
            b_f = b_f + torch.stack([fi.aggregation.bias, torch.zeros_like(bt), torch.zeros_like(bt)])
            p = F_.typed_linear(input_features, w_f, b_f, g.type_bounds)
            return _gather_scatter(p, g, None, 1.0, None, g.dv_inv)                        # :225,:233-234
        xp = F_.typed_linear(input_features, wt.unsqueeze(0), bt.unsqueeze(0), None)       # :224
        ef = fi(xp)                                                                        # :225
        return _ScatterMeanFn.apply(ef, g, g.dv_inv)                                       # :233-234


class HGCNLayer(nn.Module):
    """out = Dv^-1/2 H De^-1 H^T Dv^-1/2 Linear(X)           (GnnLayers.py:142-153)."""

    def __init__(self, device, dataset, input_dimension: int, output_dimension: int):
        super().__init__()
        self.device = device
        self.dataset = dataset
        graph: PpsHyperGraph = dataset.graph
        self.graph = graph
        self.Dv_neg_1_slash_2 = graph.dv_inv_sqrt.view(-1, 1)    # GnnLayers.py:133
        self.De_neg_1 = graph.EdgeDegrees.pow(-1)                # :134  (== 1/3)
        self._alpha = float(1.0 / 3.0)
        self._bwd_scale = graph.dv_inv_sqrt * torch.tensor(1.0 / 3.0, dtype=torch.float32, device=graph.dv_inv_sqrt.device)
        self.feature_transform = nn.Linear(input_dimension, output_dimension)

    def forward(self, input_features: Tensor) -> Tensor:
        _lib.require_cuda(input_features)
        g = self.graph
        h = F_.typed_linear(input_features, self.feature_transform.weight.unsqueeze(0),
                            self.feature_transform.bias.unsqueeze(0), None)
        return _gather_scatter(h, g, g.dv_inv_sqrt, self._alpha, self._bwd_scale, g.dv_inv_sqrt)


class GCNLayer(nn.Module):
    """out = D^-1/2 A D^-1/2 Linear(X) over the pairwise graph of Pps2DGraph
    (/root/reference/Models/GnnLayers.py:9-45; the Linear first or last exactly as :33-43).
    A x is computed from the hypergraph incidence without materialising A: for a node r,
    (A x)[r] = [self connection] x[r] + sum over its interactions of the two OTHER nodes' rows
    (`ihg_two_hop_reduce` with the own term switched off), which equals the coalesced adjacency
    product because duplicate pairs are summed either way."""

    def __init__(self, device, dataset, input_dimension: int, output_dimension: int):
        super().__init__()
        self.device = device
        self.dataset = dataset
        self.input_dimension = input_dimension
        self.output_dimension = output_dimension
        graph2d: Pps2DGraph = dataset.graph2d                          # GnnLayers.py:23
        self.graph2d = graph2d
        self.Dv_neg_1_slash_2 = graph2d.dv_inv_sqrt.view(-1, 1)        # :24
        self.feature_transform = nn.Linear(input_dimension, output_dimension)

    def _propagate(self, h: Tensor) -> Tensor:
        g2 = self.graph2d
        if g2.pair_plan is not None:            # graph_only_* completeness / flag weights: explicit pair CSR
            return _PairAdjFn.apply(h, g2)
        own = (0.0, 1.0 if g2.use_self_connection else 0.0)
        return _TwoHopFn.apply(h, g2.hyper, g2.dv_inv_sqrt, 1.0, g2.dv_inv_sqrt, own)

    def forward(self, input_features: Tensor) -> Tensor:
        _lib.require_cuda(input_features)
        w, b = self.feature_transform.weight.unsqueeze(0), self.feature_transform.bias.unsqueeze(0)
        if self.input_dimension >= self.output_dimension:              # :33-38
            return self._propagate(F_.typed_linear(input_features, w, b, None))
        return F_.typed_linear(self._propagate(input_features), w, b, None)   # :39-43


# --------------------------------------------------------------------------------------
# prediction
# --------------------------------------------------------------------------------------
class HemPredictionLayer(nn.Module):
    """score = sum_D item * (lambda*query + (1-lambda)*user) + bias[item]  (PredictionLayers.py:21-44);
    with Gs.Prediction.use_cosine_similarity the dot product becomes cosine_similarity(item, m) (:38-40)."""

    def __init__(self, feature_dimension: int, lambda_muq: float, item_count: int):
        super().__init__()
        self.feature_dimension = feature_dimension
        self.lambda_muq = lambda_muq
        self.items_bias = Parameter(Tensor(item_count))
        init.normal_(self.items_bias)

    def forward(self, user_feature: Optional[Tensor], query_feature: Tensor, item_feature: Tensor,
                item_indices: Optional[Tensor] = None) -> Tensor:
        return F_.hem_score(user_feature, query_feature, item_feature, self.items_bias, item_indices,
                            self.lambda_muq, cosine=bool(Gs.Prediction.use_cosine_similarity))   # :38-43
