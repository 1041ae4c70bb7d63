"""Raw kernel wrappers and the custom autograd Functions built on them.

Everything here calls the C ABI of libihgnn_b200.so through `_lib`; there is no torch-op
fallback for any of the hot-path computations.  torch supplies device memory (the caching
allocator), the current stream and the autograd tape.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from . import _lib
from .graph import INT64_MAX, CsrPlan

_F32 = torch.float32


def _empty(shape, like: torch.Tensor) -> torch.Tensor:
    return torch.empty(shape, dtype=_F32, device=like.device)


def _ws(nbytes: int, like: torch.Tensor) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=like.device)


# --------------------------------------------------------------------------------------
# raw ops (no autograd)
# --------------------------------------------------------------------------------------
def segment_reduce(plan: CsrPlan, src: torch.Tensor, dim: int, *, src_row_mul: int = 1,
                   bounds: Tuple[int, int] = (INT64_MAX, INT64_MAX),
                   src_scale: Optional[torch.Tensor] = None,
                   row_scale: Optional[torch.Tensor] = None,
                   out: Optional[torch.Tensor] = None,
                   row_slot: Optional[torch.Tensor] = None,
                   init: Optional[torch.Tensor] = None, accumulate: bool = False,
                   l2_source: bool = False) -> torch.Tensor:
    """out[r] = row_scale[r] * (init[r] + sum_{j in row r} src_scale[col j] * src[col[j]*mul + slot(r)]);
    accumulate: out[r] = init[r] + row_scale[r] * sum, empty rows untouched (init may be out)."""
    _lib.require_cuda(src, src_scale, row_scale, out, row_slot, init)
    if init is not None:
        init = _lib.rows_f32(init)
    if src_row_mul == 1:
        src = _lib.rows_f32(src)
        src_ld = _lib.ld(src)
    else:
        assert src.is_contiguous() and src.dtype == _F32
        src_ld = dim
    if out is None:
        out = _empty((plan.n_rows, dim), src)
    _lib.call("ihg_segment_reduce", plan.ref(), _lib.ptr(src), src_ld, src_row_mul, bounds[0],
              bounds[1], _lib.ptr(row_slot), _lib.ptr(init), _lib.ld(init) if init is not None else 0,
              _lib.ptr(src_scale), _lib.ptr(row_scale), _lib.ptr(plan.partial(dim)),
              _lib.ptr(out), _lib.ld(out), dim, (1 if accumulate else 0) | (2 if l2_source else 0), _lib.stream_ptr(),
              tag="segment_reduce",
              algo_bytes=plan.nnz * (4 + 4 * dim) + plan.n_rows * (4 * dim + 16))
    return out


def two_hop_enabled(n_rows: int, dim: int) -> bool:
    """Whether the order-1 node -> hyperedge -> node round trip runs as ONE two-hop pass over the
    node table (2 row gathers per incidence, no [E,dim] intermediate) instead of gather-sum +
    segmented reduce.  Measured on B200 (profiles/r01_bench_twohop_*.json): 0.33 vs 0.375 ms per
    round trip at the amazon-full shape (node table L2-resident) and 1.81 vs 2.32 ms at the cikm
    shape (256 MB table, NOT L2-resident: the Zipf head still hits), so it is the default;
    IHG_TWO_HOP=0 selects the two-kernel form."""
    return os.environ.get("IHG_TWO_HOP", "1") != "0"


def two_hop_reduce(plan: CsrPlan, nbr: torch.Tensor, src: torch.Tensor, *,
                   node_scale: Optional[torch.Tensor] = None, alpha: float = 1.0,
                   row_scale: Optional[torch.Tensor] = None,
                   out: Optional[torch.Tensor] = None, own=(1.0, 0.0)) -> torch.Tensor:
    """out[r] = row_scale[r] * alpha * sum_{e contains r} sum_{n in e} node_scale[n] * src[n].
    `own` = (per incidence, constant) weight of the row's own term: (1, 0) hypergraph round trip,
    (0, 0) / (0, 1) the pairwise adjacency of Pps2DGraph without / with self connections."""
    _lib.require_cuda(src, nbr, node_scale, row_scale, out)
    src = _lib.rows_f32(src)
    dim = int(src.shape[1])
    assert src.shape[0] == plan.n_rows, (src.shape, plan.n_rows)
    if out is None:
        out = _empty((plan.n_rows, dim), src)
    _lib.call("ihg_two_hop_reduce", plan.ref(), _lib.ptr(nbr), _lib.ptr(src), _lib.ld(src),
              _lib.ptr(node_scale), float(alpha), float(own[0]), float(own[1]), _lib.ptr(row_scale),
              _lib.ptr(plan.partial(dim)), _lib.ptr(out), _lib.ld(out), dim, _lib.stream_ptr(), tag="two_hop_reduce",
              # what the pair it replaces must move: gather-sum E(12+16d) + segmented reduce
              algo_bytes=(plan.nnz // 3) * (12 + 16 * dim) + plan.nnz * (4 + 4 * dim) + plan.n_rows * (4 * dim + 16))
    return out


class OutRoute:
    """Destinations of consecutive row ranges of a routed reduction (`ihg_*_routed`): range k = rows
    [starts[k], starts[k+1]) goes to device address bases[k] (+ row stride `ld` floats).  Host arrays,
    built once per call site; `keep` holds the tensors the addresses point into."""

    def __init__(self, starts, bases, ld: int, keep=()):
        import ctypes
        n = len(bases)
        assert len(starts) == n + 1 and 1 <= n <= 16
        self.n, self.ld, self.keep = n, int(ld), tuple(keep)
        self.starts = (ctypes.c_int64 * (n + 1))(*[int(x) for x in starts])
        self.bases = (ctypes.c_void_p * n)(*[int(b) if b else None for b in bases])

    def shifted(self, byte_offset: int) -> "OutRoute":
        """The same ranges `byte_offset` bytes further (a column block of the destination rows)."""
        r = OutRoute.__new__(OutRoute)
        r.n, r.ld, r.keep, r.starts = self.n, self.ld, self.keep, self.starts
        import ctypes
        r.bases = (ctypes.c_void_p * self.n)(*[(b + byte_offset) if b else None for b in self.bases])
        return r


def segment_reduce_routed(plan: CsrPlan, src: torch.Tensor, dim: int, route: OutRoute, *, src_row_mul: int = 1,
                          bounds: Tuple[int, int] = (INT64_MAX, INT64_MAX),
                          row_slot: Optional[torch.Tensor] = None) -> None:
    """`segment_reduce` (no init / scales) whose result rows go where `route` says."""
    _lib.require_cuda(src, row_slot)
    if src_row_mul == 1:
        src = _lib.rows_f32(src)
        src_ld = _lib.ld(src)
    else:
        assert src.is_contiguous() and src.dtype == _F32
        src_ld = dim
    _lib.call("ihg_segment_reduce_routed", plan.ref(), _lib.ptr(src), src_ld, src_row_mul, bounds[0], bounds[1],
              _lib.ptr(row_slot), _lib.ptr(plan.partial(dim)), route.starts, route.bases, route.n, route.ld, dim,
              _lib.stream_ptr(), tag="segment_reduce",
              algo_bytes=plan.nnz * (4 + 4 * dim) + plan.n_rows * (4 * dim + 16))


def two_hop_reduce_routed(plan: CsrPlan, nbr: torch.Tensor, src: torch.Tensor, route: OutRoute, *,
                          node_scale: Optional[torch.Tensor] = None, alpha: float = 1.0, own=(1.0, 0.0)) -> None:
    """`two_hop_reduce` (no row scale) whose result rows go where `route` says."""
    _lib.require_cuda(src, nbr, node_scale)
    src = _lib.rows_f32(src)
    dim = int(src.shape[1])
    assert src.shape[0] == plan.n_rows, (src.shape, plan.n_rows)
    _lib.call("ihg_two_hop_reduce_routed", plan.ref(), _lib.ptr(nbr), _lib.ptr(src), _lib.ld(src),
              _lib.ptr(node_scale), float(alpha), float(own[0]), float(own[1]), _lib.ptr(plan.partial(dim)),
              route.starts, route.bases, route.n, route.ld, dim, _lib.stream_ptr(), tag="two_hop_reduce",
              algo_bytes=(plan.nnz // 3) * (12 + 16 * dim) + plan.nnz * (4 + 4 * dim) + plan.n_rows * (4 * dim + 16))


def edge_gather_sum(src: torch.Tensor, i3: torch.Tensor, *, node_scale: Optional[torch.Tensor] = None,
                    alpha: float = 1.0, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[e] = alpha * sum_s node_scale[i3[e,s]] * src[i3[e,s]] (+ bias)."""
    _lib.require_cuda(src, i3, node_scale, bias)
    src = _lib.rows_f32(src)
    E, dim = int(i3.shape[0]), int(src.shape[1])
    out = _empty((E, dim), src)
    _lib.call("ihg_edge_gather_sum", _lib.ptr(src), _lib.ld(src), _lib.ptr(node_scale), float(alpha),
              _lib.ptr(bias), _lib.ptr(i3), E, _lib.ptr(out), dim, dim, _lib.stream_ptr(),
              tag="edge_gather_sum", algo_bytes=E * (12 + 16 * dim + (12 if node_scale is not None else 0)))
    return out


def node_linear(x: torch.Tensor, w: torch.Tensor, *, transpose_w: bool = False,
                bias: Optional[torch.Tensor] = None, addend: Optional[torch.Tensor] = None,
                bounds: Optional[Tuple[int, int]] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Typed Linear.  w: [T, n_out, n_in] (transpose_w False) -> y = x W[t]^T + bias[t] + addend;
    transpose_w True -> y = x W[t] (n_out = w.shape[2])."""
    _lib.require_cuda(x, w, bias, addend)
    x = _lib.rows_f32(x)
    w = w.contiguous()
    assert w.dim() == 3 and w.dtype == _F32
    T, r, c = (int(s) for s in w.shape)
    n_out, n_in = (c, r) if transpose_w else (r, c)
    assert x.shape[1] == n_in, (x.shape, w.shape, transpose_w)
    n_rows = int(x.shape[0])
    if T == 1:
        b0 = b1 = n_rows
    else:
        assert T == 3 and bounds is not None
        b0, b1 = bounds
    if bias is not None:
        bias = bias.contiguous()
    if addend is not None:
        addend = _lib.rows_f32(addend)
    y = _empty((n_rows, n_out), x) if out is None else out
    assert tuple(y.shape) == (n_rows, n_out) and y.dtype == _F32 and y.stride(1) == 1 and y.data_ptr() % 16 == 0
    _lib.call("ihg_node_linear", _lib.ptr(x), _lib.ld(x), _lib.ptr(w), T, n_out, n_in,
              1 if transpose_w else 0, _lib.ptr(bias), _lib.ptr(addend),
              _lib.ld(addend) if addend is not None else 0, n_rows, b0, b1, _lib.ptr(y), _lib.ld(y),
              _lib.stream_ptr(), tag="node_linear",
              algo_bytes=n_rows * 4 * (n_in + n_out + (n_out if addend is not None else 0)))
    return y


def node_linear_wgrad(dy: torch.Tensor, x: torch.Tensor, n_types: int,
                      bounds: Optional[Tuple[int, int]], want_bias: bool):
    """dw[t] = sum_{r in t} dy[r]^T x[r] -> [T, n_out, n_in]; db[t] = sum dy[r] -> [T, n_out]."""
    dy = _lib.rows_f32(dy)
    x = _lib.rows_f32(x)
    n_rows, n_out, n_in = int(x.shape[0]), int(dy.shape[1]), int(x.shape[1])
    b0, b1 = (n_rows, n_rows) if n_types == 1 else bounds
    dw = _empty((n_types, n_out, n_in), x)
    db = _empty((n_types, n_out), x) if want_bias else None
    ws_bytes = _lib.lib().ihg_node_linear_wgrad_workspace_bytes(n_types, n_out, n_in)
    ws = _ws(ws_bytes, x)
    _lib.call("ihg_node_linear_wgrad", _lib.ptr(dy), _lib.ld(dy), _lib.ptr(x), _lib.ld(x), n_rows,
              b0, b1, n_types, n_out, n_in, _lib.ptr(dw), _lib.ptr(db), _lib.ptr(ws), ws_bytes,
              _lib.stream_ptr(), tag="node_linear_wgrad", algo_bytes=n_rows * 4 * (n_in + n_out))
    return dw, db


def gather_rows_raw(table: torch.Tensor, idx: torch.Tensor, offset: int = 0) -> torch.Tensor:
    """out[b] = table[idx[b] + offset] (no autograd); idx int64 on the device."""
    _lib.require_cuda(table, idx)
    table = _lib.rows_f32(table)
    B, d = int(idx.numel()), int(table.shape[1])
    out = _empty((B, d), table)
    if B:
        _lib.call("ihg_gather_rows", _lib.ptr(table), _lib.ld(table), _lib.ptr(idx), offset, B,
                  _lib.ptr(out), d, d, _lib.stream_ptr(), tag="gather_rows", algo_bytes=B * (8 + 8 * d))
    return out


def copy_rows_raw(src: torch.Tensor, dst: torch.Tensor) -> None:
    """dst[r] = src[r] for 2-D row-major tensors with equal shapes (128-bit vectorised)."""
    _lib.require_cuda(src, dst)
    src = _lib.rows_f32(src)
    n, d = int(src.shape[0]), int(src.shape[1])
    if n:
        _lib.call("ihg_copy_rows", _lib.ptr(src), _lib.ld(src), _lib.ptr(dst), _lib.ld(dst), n, d,
                  _lib.stream_ptr(), tag="copy_rows", algo_bytes=n * 8 * d)


# --------------------------------------------------------------------------------------
# autograd Functions
# --------------------------------------------------------------------------------------
class TypedLinearFn(torch.autograd.Function):
    """y = x W[type(row)]^T + b[type(row)];  w [T, n_out, n_in], b [T, n_out] or None."""

    @staticmethod
    def forward(ctx, x, w, b, bounds, out=None):
        ctx.bounds = bounds
        ctx.has_bias = b is not None
        x = _lib.rows_f32(x)
        w = w.contiguous()
        ctx.save_for_backward(x, w)
        return node_linear(x, w, bias=b, bounds=bounds, out=out.tensor if out is not None else None)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = _lib.rows_f32(dy)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = node_linear(dy, w, transpose_w=True, bounds=ctx.bounds)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw, db = node_linear_wgrad(dy, x, int(w.shape[0]), ctx.bounds, ctx.has_bias)
        return dx, dw, db, None, None


class OutBuffer:
    """Wraps a preallocated result tensor handed to an autograd Function: a plain object, so autograd
    does not treat the buffer as an input of the Function (the result is a fresh view of it)."""

    def __init__(self, tensor: torch.Tensor):
        self.tensor = tensor


def typed_linear(x, w, b=None, bounds=None, out: Optional[torch.Tensor] = None):
    """`out`: write the result into this (non-differentiable) buffer instead of a fresh tensor -- the
    multi-GPU layers project straight into the head of the table their halo exchange completes."""
    return TypedLinearFn.apply(x, w, b, bounds, OutBuffer(out) if out is not None else None)


class EmbedAllFn(torch.autograd.Function):
    """EmbeddingLayer.forward(None, None, None) fused with RawGnn's cat
    (/root/reference/Models/EmbeddingLayers.py:63-81, Models/RawGnn.py:112):
        X = [ weight_user[1:] ; bag_mean(weight_vocab) ; weight_item[1:] ]      [N, d]
    Backward: dense table gradients with the padding row 0 zero; the vocabulary gradient is a
    segmented reduce over the word -> query transpose (deterministic)."""

    @staticmethod
    def forward(ctx, w_user, w_vocab, w_item, tables):
        _lib.require_cuda(w_user, w_vocab, w_item)
        U, Q, I = tables.user_count, tables.query_count, tables.item_count
        d = int(w_user.shape[1])
        w_user, w_vocab, w_item = (_lib.rows_f32(t) for t in (w_user, w_vocab, w_item))
        x = _empty((U + Q + I, d), w_user)
        st = _lib.stream_ptr()
        _lib.call("ihg_copy_rows", _lib.ptr(w_user) + 4 * _lib.ld(w_user), _lib.ld(w_user),
                  _lib.ptr(x), d, U, d, st)
        segment_reduce(tables.bag_plan, w_vocab, d, row_scale=tables.bag_inv_len, out=x[U:U + Q])
        _lib.call("ihg_copy_rows", _lib.ptr(w_item) + 4 * _lib.ld(w_item), _lib.ld(w_item),
                  _lib.ptr(x) + 4 * d * (U + Q), d, I, d, st)
        ctx.tables = tables
        ctx.shapes = (tuple(w_user.shape), tuple(w_vocab.shape), tuple(w_item.shape))
        return x

    @staticmethod
    def backward(ctx, dx):
        t = ctx.tables
        U, Q, I = t.user_count, t.query_count, t.item_count
        dx = _lib.rows_f32(dx)
        d = int(dx.shape[1])
        ldx = _lib.ld(dx)
        st = _lib.stream_ptr()
        su, sv, si = ctx.shapes
        dwu = dwv = dwi = None
        if ctx.needs_input_grad[0]:
            dwu = _empty(su, dx)
            dwu[0].zero_()                                   # padding_idx=0 row: zero gradient
            _lib.call("ihg_copy_rows", _lib.ptr(dx), ldx, _lib.ptr(dwu) + 4 * d, d, U, d, st)
        if ctx.needs_input_grad[1]:
            # dW_vocab[w] = sum over occurrences (q, w) of dX[U+q] / len(q)
            dwv = segment_reduce(t.word_plan, dx[U:U + Q], d, src_scale=t.bag_inv_len)
        if ctx.needs_input_grad[2]:
            dwi = _empty(si, dx)
            dwi[0].zero_()
            _lib.call("ihg_copy_rows", _lib.ptr(dx) + 4 * ldx * (U + Q), ldx,
                      _lib.ptr(dwi) + 4 * d, d, I, d, st)
        return dwu, dwv, dwi, None


class GatherRowsFn(torch.autograd.Function):
    """out[b] = table[idx[b] + offset]; backward is a deterministic scatter-add into a dense
    gradient (the reference's `index` / `embedding` backward)."""

    @staticmethod
    def forward(ctx, table, idx, offset: int):
        _lib.require_cuda(table, idx)
        table = _lib.rows_f32(table)
        idx = idx.to(torch.int64).contiguous()
        B, d = int(idx.numel()), int(table.shape[1])
        out = _empty((B, d), table)
        _lib.call("ihg_gather_rows", _lib.ptr(table), _lib.ld(table), _lib.ptr(idx), offset, B,
                  _lib.ptr(out), d, d, _lib.stream_ptr())
        ctx.save_for_backward(idx)
        ctx.offset, ctx.shape = offset, tuple(table.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        g = _lib.rows_f32(g)
        d = int(g.shape[1])
        dt = torch.zeros(ctx.shape, dtype=_F32, device=g.device)
        _lib.call("ihg_scatter_add_rows", _lib.ptr(g), _lib.ld(g), _lib.ptr(idx), ctx.offset,
                  int(idx.numel()), _lib.ptr(dt), d, d, _lib.stream_ptr())
        return dt, None, None


def gather_rows(table, idx, offset: int = 0):
    return GatherRowsFn.apply(table, idx, offset)


class GatherRowsMultiFn(torch.autograd.Function):
    """Several row selections out of ONE table (the user / query / item rows of a batch,
    RawGnn.py:128-133): outs[j][b] = table[idx_j[b] + offset_j].  Backward builds the dense table
    gradient once and scatter-adds every selection into it in order (deterministic), instead of one
    dense gradient per selection plus autograd's accumulation adds."""

    @staticmethod
    def forward(ctx, table, offsets, *idxs):
        _lib.require_cuda(table, *idxs)
        table = _lib.rows_f32(table)
        idxs = tuple(i.to(torch.int64).contiguous() for i in idxs)
        d = int(table.shape[1])
        outs = []
        for idx, off in zip(idxs, offsets):
            B = int(idx.numel())
            out = _empty((B, d), table)
            if B:
                _lib.call("ihg_gather_rows", _lib.ptr(table), _lib.ld(table), _lib.ptr(idx), int(off), B,
                          _lib.ptr(out), d, d, _lib.stream_ptr())
            outs.append(out)
        ctx.save_for_backward(*idxs)
        ctx.offsets, ctx.shape = tuple(int(o) for o in offsets), tuple(table.shape)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        idxs = ctx.saved_tensors
        dt = torch.zeros(ctx.shape, dtype=_F32, device=idxs[0].device)
        d = ctx.shape[1]
        for g, idx, off in zip(gs, idxs, ctx.offsets):
            if g is None or idx.numel() == 0:
                continue
            g = _lib.rows_f32(g)
            _lib.call("ihg_scatter_add_rows", _lib.ptr(g), _lib.ld(g), _lib.ptr(idx), off,
                      int(idx.numel()), _lib.ptr(dt), d, d, _lib.stream_ptr())
        return (dt, None) + (None,) * len(idxs)


def gather_rows_multi(table, idxs, offsets):
    return GatherRowsMultiFn.apply(table, tuple(offsets), *idxs)


class TapRowsFn(torch.autograd.Function):
    """`table` passed through unchanged PLUS several row selections out of it (the batch rows RawGnn.py:128-133
    takes out of every layer's output while the same output feeds the next layer).  Backward adds the few
    selected-row gradients INTO the dense gradient arriving for the pass-through output (in place, fixed order)
    instead of building a second dense [N, d] gradient and letting autograd add the two: per layer output that
    saves one [N, d] zero-fill and one three-operand dense add."""

    @staticmethod
    def forward(ctx, table, offsets, *idxs):
        _lib.require_cuda(table, *idxs)
        src = _lib.rows_f32(table)
        idxs = tuple(i.to(torch.int64).contiguous() for i in idxs)
        d = int(src.shape[1])
        outs = []
        for idx, off in zip(idxs, offsets):
            B = int(idx.numel())
            out = _empty((B, d), src)
            if B:
                _lib.call("ihg_gather_rows", _lib.ptr(src), _lib.ld(src), _lib.ptr(idx), int(off), B,
                          _lib.ptr(out), d, d, _lib.stream_ptr())
            outs.append(out)
        ctx.save_for_backward(*idxs)
        ctx.offsets, ctx.shape = tuple(int(o) for o in offsets), tuple(table.shape)
        return (table.view_as(table),) + tuple(outs)

    @staticmethod
    def backward(ctx, g_table, *gs):
        idxs = ctx.saved_tensors
        d = ctx.shape[1]
        if g_table is None:
            dt = torch.zeros(ctx.shape, dtype=_F32, device=idxs[0].device)
        elif g_table.is_contiguous() and g_table.dtype == _F32 and g_table.data_ptr() % 16 == 0:
            dt = g_table                       # a temporary produced by the consumer's backward: accumulate in place
        else:
            dt = g_table.contiguous().to(_F32).clone()
        for g, idx, off in zip(gs, idxs, ctx.offsets):
            if g is None or idx.numel() == 0:
                continue
            g = _lib.rows_f32(g)
            _lib.call("ihg_scatter_add_rows", _lib.ptr(g), _lib.ld(g), _lib.ptr(idx), off,
                      int(idx.numel()), _lib.ptr(dt), d, d, _lib.stream_ptr())
        return (dt, None) + (None,) * len(idxs)


def tap_rows(table, idxs, offsets):
    """-> (table, [table[idx_j + offset_j] for j]) with the fused backward of `TapRowsFn`."""
    out = TapRowsFn.apply(table, tuple(offsets), *idxs)
    return out[0], out[1:]


class HemScoreFn(torch.autograd.Function):
    """HemPredictionLayer.forward, dot-product branch
    (/root/reference/Models/PredictionLayers.py:21-44)."""

    @staticmethod
    def forward(ctx, user_f, query_f, item_f, items_bias, item_idx, lam: float, cosine: bool = False):
        _lib.require_cuda(user_f, query_f, item_f, items_bias, item_idx)
        query_f, item_f = _lib.rows_f32(query_f), _lib.rows_f32(item_f)
        if user_f is not None:
            user_f = _lib.rows_f32(user_f)
        if item_idx is not None:
            item_idx = item_idx.to(torch.int64).contiguous()
        items_bias = items_bias.contiguous()
        B, D = int(item_f.shape[0]), int(item_f.shape[1])
        score = torch.empty(B, dtype=_F32, device=item_f.device)
        norms = torch.empty((B, 3), dtype=_F32, device=item_f.device) if cosine else None   # (item.m, |item|, |m|)
        _lib.call("ihg_hem_score_fwd", _lib.ptr(user_f), _lib.ld(user_f) if user_f is not None else 0,
                  _lib.ptr(query_f), _lib.ld(query_f), _lib.ptr(item_f), _lib.ld(item_f),
                  _lib.ptr(items_bias), _lib.ptr(item_idx), float(lam), B, D, _lib.ptr(score),
                  1 if cosine else 0, _lib.ptr(norms), _lib.stream_ptr())
        ctx.norms = norms
        ctx.lam, ctx.n_items = float(lam), int(items_bias.numel())
        ctx.has_user, ctx.has_idx = user_f is not None, item_idx is not None
        ctx.save_for_backward(*(t for t in (user_f, query_f, item_f, item_idx) if t is not None))
        return score

    @staticmethod
    def backward(ctx, dscore):
        saved = list(ctx.saved_tensors)
        user_f = saved.pop(0) if ctx.has_user else None
        query_f, item_f = saved.pop(0), saved.pop(0)
        item_idx = saved.pop(0) if ctx.has_idx else None
        dscore = dscore.contiguous()
        B, D = int(item_f.shape[0]), int(item_f.shape[1])
        need = ctx.needs_input_grad
        d_user = _empty((B, D), item_f) if (ctx.has_user and need[0]) else None
        d_query = _empty((B, D), item_f) if need[1] else None
        d_item = _empty((B, D), item_f) if need[2] else None
        d_bias = torch.empty(ctx.n_items, dtype=_F32, device=item_f.device) if need[3] else None
        ws_bytes = _lib.lib().ihg_hem_score_bwd_workspace_bytes(ctx.n_items) if (need[3] and item_idx is not None) else 0
        ws = torch.empty(ws_bytes // 8, dtype=torch.int64, device=item_f.device) if ws_bytes else None
        _lib.call("ihg_hem_score_bwd", _lib.ptr(dscore), _lib.ptr(user_f),
                  _lib.ld(user_f) if user_f is not None else 0, _lib.ptr(query_f), _lib.ld(query_f),
                  _lib.ptr(item_f), _lib.ld(item_f), _lib.ptr(item_idx), ctx.lam, B, D,
                  _lib.ptr(d_user), _lib.ptr(d_query), _lib.ptr(d_item), _lib.ptr(d_bias),
                  ctx.n_items, _lib.ptr(ws), ws_bytes, _lib.ptr(ctx.norms), _lib.stream_ptr())
        return d_user, d_query, d_item, d_bias, None, None, None


def hem_score(user_f, query_f, item_f, items_bias, item_idx, lam, cosine: bool = False):
    return HemScoreFn.apply(user_f, query_f, item_f, items_bias, item_idx, lam, bool(cosine))


def rank_topk(features: torch.Tensor, users: Optional[torch.Tensor], queries: torch.Tensor,
              items_bias: torch.Tensor, lam: float, *, query_row0: int, item_row0: int, item_count: int,
              candidates: Optional[torch.Tensor] = None, k: int = 10, cosine: bool = False):
    """Batched inference ranking (no autograd): for every (user, query) the k best of the candidate
    items (`candidates` int64 [B, C]; None = all items) under the HEM score
    (/root/reference/Models/PredictionLayers.py:21-44), i.e. what
    `torch.sort(model(users, queries, None), descending=True)[1][:10]`
    (/root/reference/Helpers/Metrics.py:60-61) yields per query.  Returns (item ids int64 [B, k],
    scores fp32 [B, k])."""
    _lib.require_cuda(features, users, queries, items_bias, candidates)
    features = _lib.rows_f32(features)
    queries = queries.to(torch.int64).contiguous()
    if users is not None:
        users = users.to(torch.int64).contiguous()
        assert users.numel() == queries.numel()
    B, D = int(queries.numel()), int(features.shape[1])
    if candidates is not None:
        candidates = candidates.to(torch.int64).contiguous()
        assert candidates.dim() == 2 and candidates.shape[0] == B, (candidates.shape, B)
        C = int(candidates.shape[1])
    else:
        C = int(item_count)
    items_bias = items_bias.detach().contiguous()
    top_items = torch.empty((B, k), dtype=torch.int64, device=features.device)
    top_scores = torch.empty((B, k), dtype=_F32, device=features.device)
    _lib.call("ihg_rank_topk", _lib.ptr(features), _lib.ld(features), _lib.ptr(users), _lib.ptr(queries), B,
              int(query_row0), _lib.ptr(candidates), C, int(item_row0), int(item_count), _lib.ptr(items_bias),
              float(lam), D, int(k), 1 if cosine else 0, _lib.ptr(top_items), _lib.ptr(top_scores), _lib.stream_ptr(),
              tag="rank_topk", algo_bytes=B * (8 * D + C * ((8 if candidates is not None else 0) + 4 * D + 4) + 12 * k))
    return top_items, top_scores
