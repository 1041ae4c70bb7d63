"""Device-resident hypergraph indices: drop-in for `Helpers.Graph.PpsHyperGraph`.

`PpsHyperGraph.from_interactions(interactions, node_count, user_count, query_count, device)`
keeps the reference signature (/root/reference/Helpers/Graph.py:94-100) and still exposes
`.Adjacency .I3 .VertexDegrees .EdgeDegrees .EdgeCount` (Graph.py:86-92), but the data is
built on the GPU by `ihg_graph_build` as sorted int32 CSR (node -> hyperedge: `rowptr`,
`col`) and CSC (hyperedge -> node: `i3`) arrays; the int64 / sparse-COO views the reference
exposed are materialised lazily, only if something reads them.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np
import torch

from . import _lib

INT64_MAX = (1 << 63) - 1
DEFAULT_CHUNK_LEN = int(os.environ.get("IHG_CHUNK_LEN", "128"))   # incidences per work item; 128 measured best (dev switch)


class CsrPlan:
    """A unit-valued CSR matrix on the device plus the deterministic load-balancing plan of
    `ihg_segment_plan_build` (rows longer than `chunk_len` are split).  Owns the tensors the
    `ihg_csr` struct points into."""

    def __init__(self, rowptr: torch.Tensor, col: torch.Tensor, chunk_len: int = DEFAULT_CHUNK_LEN,
                 drop_empty_rows: bool = False):
        _lib.require_cuda(rowptr, col)
        assert rowptr.dtype == torch.int32 and col.dtype == torch.int32
        self.rowptr = rowptr.contiguous()
        self.col = col.contiguous()
        self.n_rows = int(rowptr.numel() - 1)
        self.nnz = int(col.numel())
        self.chunk_len = int(chunk_len)
        dev = rowptr.device
        cap_extra = self.nnz // self.chunk_len
        cap_seg = self.n_rows + cap_extra + 1
        seg = torch.empty((cap_seg, 4), dtype=torch.int32, device=dev)
        split_row = torch.empty(cap_extra + 1, dtype=torch.int32, device=dev)
        split_ptr = torch.empty(cap_extra + 2, dtype=torch.int32, device=dev)
        counts = torch.zeros(3, dtype=torch.int64, device=dev)
        ws_bytes = _lib.lib().ihg_segment_plan_workspace_bytes(self.n_rows)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.call("ihg_segment_plan_build", _lib.ptr(self.rowptr), self.n_rows, self.chunk_len,
                  _lib.ptr(seg), _lib.ptr(split_row), _lib.ptr(split_ptr), _lib.ptr(counts),
                  _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
        n_seg, n_split, n_part = (int(x) for x in counts.tolist())   # one-time sync at build
        self.n_seg, self.n_split, self.n_part = n_seg, n_split, n_part
        self.seg = seg[:n_seg]                 # [n_seg, 4] = (begin, end, row, partial slot | -1)
        if drop_empty_rows:                    # accumulate-mode plans: rows without incidences are never touched
            self.seg = self.seg[self.seg[:, 1] > self.seg[:, 0]].contiguous()
            n_seg = int(self.seg.shape[0])
            self.n_seg = n_seg
            if n_seg == 0:
                self.seg = torch.zeros((1, 4), dtype=torch.int32, device=dev)
        self.split_row = split_row[:max(n_split, 1)]
        self.split_ptr = split_ptr[:n_split + 1]
        self.struct = _lib.IhgCsr(
            n_rows=self.n_rows, nnz=self.nnz, rowptr=self.rowptr.data_ptr(),
            col=self.col.data_ptr() if self.nnz else None, chunk_len=self.chunk_len,
            n_seg=n_seg, n_split=n_split, n_part=n_part, seg=self.seg.data_ptr(),
            split_row=self.split_row.data_ptr(), split_ptr=self.split_ptr.data_ptr())
        self._partial = {}
        self._nbr = None

    def partial(self, dim: int) -> Optional[torch.Tensor]:
        """Scratch for the partial sums of split rows (cached per feature dimension)."""
        if self.n_part == 0:
            return None
        buf = self._partial.get(dim)
        if buf is None:
            buf = torch.empty(self.n_part * dim, dtype=torch.float32, device=self.rowptr.device)
            self._partial[dim] = buf
        return buf

    def ref(self):
        return ctypes.byref(self.struct)

    def two_hop_nbr(self, i3: torch.Tensor, bounds=(INT64_MAX, INT64_MAX),
                    row_slot: Optional[torch.Tensor] = None) -> torch.Tensor:
        """int32 [nnz, 2]: for every incidence (row r, hyperedge col[j]) the two OTHER nodes of that
        hyperedge (`ihg_two_hop_index_build`); built once per plan, the graph is static."""
        if self._nbr is None:
            _lib.require_cuda(i3, row_slot)
            nbr = torch.empty((max(self.nnz, 1), 2), dtype=torch.int32, device=self.rowptr.device)
            _lib.call("ihg_two_hop_index_build", self.ref(), _lib.ptr(i3), bounds[0], bounds[1],
                      _lib.ptr(row_slot), _lib.ptr(nbr), _lib.stream_ptr())
            self._nbr = nbr
        return self._nbr


def csr_from_keys(keys: torch.Tensor, num_keys: int, values: Optional[torch.Tensor] = None):
    """Stable CSR of an int32 key array on the device: (rowptr[num_keys+1], perm[n],
    values[perm] or None).  Raises on out-of-range keys."""
    _lib.require_cuda(keys)
    keys = keys.to(torch.int32).contiguous()
    n = int(keys.numel())
    dev = keys.device
    rowptr = torch.empty(num_keys + 1, dtype=torch.int32, device=dev)
    perm = torch.empty(n, dtype=torch.int32, device=dev)
    out_values = None
    if values is not None:
        values = values.to(torch.int32).contiguous()
        out_values = torch.empty(n, dtype=torch.int32, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = _lib.lib().ihg_csr_from_keys_workspace_bytes(n, num_keys)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.call("ihg_csr_from_keys", _lib.ptr(keys), _lib.ptr(values), n, num_keys, _lib.ptr(rowptr),
              _lib.ptr(perm), _lib.ptr(out_values), _lib.ptr(flag), _lib.ptr(ws), ws_bytes,
              _lib.stream_ptr())
    if int(flag.item()) != 0:
        raise ValueError("csr_from_keys: key out of range")
    return rowptr, perm, out_values


class PpsGraph:
    """Base class, as in Helpers/Graph.py:7-9."""

    def __init__(self):
        pass


class PpsHyperGraph(PpsGraph):
    """3-uniform hypergraph of positive (user, query, item) interactions.

    Reference attributes (Graph.py:86-92): Adjacency, I3, VertexDegrees, EdgeDegrees, EdgeCount.
    Device-native attributes: i3 int32 [E,3], rowptr int32 [N+1], col int32 [3E],
    dv_inv fp32 [N] (= VertexDegrees^-1, GnnLayers.py:187), dv_inv_sqrt fp32 [N]
    (= VertexDegrees^-1/2, GnnLayers.py:133), plan (CsrPlan), user_count / query_count /
    item_count / node_count.
    """

    def __init__(self):
        super().__init__()
        self._adjacency = None
        self._I3 = None

    # ---- reference entry point ---------------------------------------------------------
    @classmethod
    def from_interactions(cls, interactions, node_count: int, user_count: int, query_count: int,
                          device) -> "PpsHyperGraph":
        """Signature of Graph.py:94-100.  `interactions` is the reference's list of
        `PosInteraction` (anything with `.uqif() -> (u, q, i, flag)`), or a sequence of
        (u, q, i[, flag]) tuples; flag <= 0 entries are skipped (Graph.py:108)."""
        rows = []
        for p in interactions:
            t = p.uqif() if hasattr(p, "uqif") else tuple(p)
            if len(t) < 4 or t[3] > 0:
                rows.append(t[:3])
        arr = np.asarray(rows, dtype=np.int64).reshape(-1, 3)
        item_count = node_count - user_count - query_count
        return cls.from_tensors(torch.from_numpy(arr[:, 0].copy()), torch.from_numpy(arr[:, 1].copy()),
                                torch.from_numpy(arr[:, 2].copy()), user_count, query_count,
                                item_count, device)

    # ---- array entry point (multi-million-edge workloads skip the Python objects) -------
    @classmethod
    def from_tensors(cls, user, query, item, user_count: int, query_count: int, item_count: int,
                     device, chunk_len: int = DEFAULT_CHUNK_LEN) -> "PpsHyperGraph":
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("ihgnn_b200.PpsHyperGraph is built on the GPU; device must be CUDA "
                               f"(got {device}) -- there is no CPU fallback")
        with torch.cuda.device(device):
            u = torch.as_tensor(user).to(device=device, dtype=torch.int64).contiguous()
            q = torch.as_tensor(query).to(device=device, dtype=torch.int64).contiguous()
            i = torch.as_tensor(item).to(device=device, dtype=torch.int64).contiguous()
            E = int(u.numel())
            assert q.numel() == E and i.numel() == E
            N = user_count + query_count + item_count
            g = cls()
            g.user_count, g.query_count, g.item_count, g.node_count = user_count, query_count, item_count, N
            g.EdgeCount = E
            g.i3 = torch.empty((E, 3), dtype=torch.int32, device=device)
            g.rowptr = torch.empty(N + 1, dtype=torch.int32, device=device)
            g.col = torch.empty(3 * E, dtype=torch.int32, device=device)
            deg = torch.empty(N, dtype=torch.float32, device=device)
            g.dv_inv = torch.empty(N, dtype=torch.float32, device=device)
            g.dv_inv_sqrt = torch.empty(N, dtype=torch.float32, device=device)
            flag = torch.zeros(1, dtype=torch.int32, device=device)
            ws_bytes = _lib.lib().ihg_graph_workspace_bytes(E, N)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
            _lib.call("ihg_graph_build", _lib.ptr(u), _lib.ptr(q), _lib.ptr(i), E, user_count,
                      query_count, item_count, _lib.ptr(g.i3), _lib.ptr(g.rowptr), _lib.ptr(g.col),
                      _lib.ptr(deg), _lib.ptr(g.dv_inv), _lib.ptr(g.dv_inv_sqrt), _lib.ptr(flag),
                      _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
            if int(flag.item()) != 0:
                raise ValueError("PpsHyperGraph: user/query/item index out of range")
            del ws
            g.VertexDegrees = deg.view(-1, 1)                                   # Graph.py:131
            g.EdgeDegrees = torch.full((E, 1), 3.0, dtype=torch.float32, device=device)  # :132
            g.plan = CsrPlan(g.rowptr, g.col, chunk_len)
        return g

    # ---- lazily materialised reference views -------------------------------------------
    @property
    def I3(self) -> torch.Tensor:
        """int64 [E,3] interaction matrix (Graph.py:129)."""
        if self._I3 is None:
            self._I3 = self.i3.to(torch.int64)
        return self._I3

    @property
    def Adjacency(self) -> torch.Tensor:
        """Coalesced sparse COO incidence matrix [N,E], unit fp32 values (Graph.py:123-128)."""
        if self._adjacency is None:
            counts = (self.rowptr[1:] - self.rowptr[:-1]).to(torch.int64)
            rows = torch.repeat_interleave(torch.arange(self.node_count, device=self.rowptr.device), counts)
            idx = torch.stack([rows, self.col.to(torch.int64)])
            vals = torch.ones(idx.shape[1], dtype=torch.float32, device=self.rowptr.device)
            self._adjacency = torch.sparse_coo_tensor(idx, vals, (self.node_count, self.EdgeCount),
                                                      is_coalesced=True)
        return self._adjacency

    @property
    def type_bounds(self):
        """(U, U+Q): node ids below U are users (slot 0), below U+Q queries (slot 1), else items."""
        return self.user_count, self.user_count + self.query_count


class Pps2DGraph(PpsGraph):
    """Pairwise user-query-item graph, drop-in for `Helpers.Graph.Pps2DGraph`
    (/root/reference/Helpers/Graph.py:12-81).

    `Gs.graph_completeness == graph_uqi` with unit flags (the reference default: ArgsParser.py:85, and
    `treat_all_1` Dataset.py:200 clamps every flag to 1): every positive interaction adds u-q, q-i, i-u in
    both directions, duplicate pairs summed by `coalesce()`; `VertexDegrees` = [self connection] + 2 x
    interactions of the node (0 stored as 1e-8 without self connections, Graph.py:35,67-68).  The adjacency
    is never stored: products with it run on the hypergraph incidence (`hyper`, the same device CSR/CSC the
    IHGNN layers use) through `ihg_two_hop_reduce`.

    The other branches of Graph.py:40-65 -- `graph_only_uq / ui / qi` (one pair per interaction, degree +1
    for its two nodes) and interaction flags above 1 (they weight the u-i pair of `graph_uqi`, :44) -- take
    the *pair form*: the directed pairs are laid out as a device CSR (`pair_plan`, rows by stable sort) and
    products run through `ihg_segment_reduce`.  Flags are integers (`PosInteraction.interaction: int`,
    SearchLog.py:193) and enter as multiplicities of the pair, which is what `coalesce()` sums them to.

    Reference attributes `Adjacency` (coalesced sparse COO) and `VertexDegrees` are available; `Adjacency`
    is materialised lazily, only if something reads it."""

    def __init__(self):
        super().__init__()
        self._adjacency = None
        self.pair_plan: Optional[CsrPlan] = None

    @classmethod
    def from_interactions(cls, interactions, node_count: int, user_count: int, query_count: int,
                          use_self_connection: bool, device) -> "Pps2DGraph":
        """Signature of Graph.py:19-26."""
        rows, flags = [], []
        for p in interactions:
            t = p.uqif() if hasattr(p, "uqif") else tuple(p)
            if len(t) < 4 or t[3] > 0:                                         # Graph.py:37
                rows.append(t[:3])
                flags.append(1 if len(t) < 4 else int(t[3]))
        arr = np.asarray(rows, dtype=np.int64).reshape(-1, 3)
        item_count = node_count - user_count - query_count
        hyper = PpsHyperGraph.from_tensors(torch.from_numpy(arr[:, 0].copy()), torch.from_numpy(arr[:, 1].copy()),
                                           torch.from_numpy(arr[:, 2].copy()), user_count, query_count,
                                           item_count, device)
        fl = np.asarray(flags, dtype=np.int64)
        return cls.from_hypergraph(hyper, use_self_connection, flags=None if (fl == 1).all() else fl)

    @classmethod
    def from_hypergraph(cls, hyper: PpsHyperGraph, use_self_connection: bool, flags=None) -> "Pps2DGraph":
        """`flags`: per-hyperedge integer interaction flags in the order of `hyper.i3` (None = all 1)."""
        from .settings import Gs, Gsv
        completeness = getattr(Gs, "graph_completeness", Gsv.graph_uqi)
        pairs = {Gsv.graph_uqi: ((0, 1), (1, 2), (2, 0)), Gsv.graph_only_uq: ((0, 1),),
                 Gsv.graph_only_ui: ((0, 2),), Gsv.graph_only_qi: ((1, 2),)}.get(completeness)
        if pairs is None:
            raise ValueError(f"unknown graph_completeness {completeness!r}")           # Graph.py:64-65
        g = cls()
        g.hyper = hyper
        g.use_self_connection = bool(use_self_connection)
        g.node_count = N = hyper.node_count
        dev = hyper.i3.device
        self_deg = 1.0 if use_self_connection else 0.0                                 # Graph.py:29
        if completeness == Gsv.graph_uqi and flags is None:
            counts = (hyper.rowptr[1:] - hyper.rowptr[:-1]).to(torch.float32)
            deg = 2.0 * counts + self_deg                                              # :45
        else:
            i3 = hyper.i3.to(torch.int64)
            a = torch.cat([i3[:, s] for s, _ in pairs] + [i3[:, t] for _, t in pairs])   # both directions
            b = torch.cat([i3[:, t] for _, t in pairs] + [i3[:, s] for s, _ in pairs])
            per_node = 2.0 if completeness == Gsv.graph_uqi else 1.0                   # :45 / :50,56,62
            touched = torch.cat([i3[:, s] for s in sorted({s for pr in pairs for s in pr})])
            deg = per_node * torch.bincount(touched, minlength=N).to(torch.float32) + self_deg
            if flags is not None and completeness == Gsv.graph_uqi:                    # :44: the u-i pair carries the flag
                fl = torch.as_tensor(np.asarray(flags), dtype=torch.int64, device=dev)
                assert fl.numel() == hyper.EdgeCount and bool((fl >= 1).all())
                ones = torch.ones_like(fl)
                mult = torch.cat([fl if set(pr) == {0, 2} else ones for pr in pairs] * 2)
                a, b = torch.repeat_interleave(a, mult), torch.repeat_interleave(b, mult)
            g._pair_rows, g._pair_cols = a, b
            rowptr, _perm, cols = csr_from_keys(a.to(torch.int32), N, values=b.to(torch.int32))
            g.pair_plan = CsrPlan(rowptr, cols)
        if not use_self_connection:
            deg = torch.where(deg == 0, torch.full_like(deg, 1e-8), deg)               # :67-68
        g.VertexDegrees = deg.view(-1, 1)                                              # :80
        g.dv_inv_sqrt = deg.pow(-0.5)                                                  # GnnLayers.py:24
        return g

    @property
    def Adjacency(self) -> torch.Tensor:
        """Coalesced sparse COO adjacency [N,N] (Graph.py:71-77), weights summed over duplicates."""
        if self._adjacency is None:
            if self.pair_plan is not None:
                rows, cols = self._pair_rows, self._pair_cols
            else:
                i3 = self.hyper.i3.to(torch.int64)
                u, q, i = i3[:, 0], i3[:, 1], i3[:, 2]
                rows = torch.cat([u, q, i, i, q, u])                          # Graph.py:42-43
                cols = torch.cat([q, i, u, q, u, i])
            if self.use_self_connection:
                eye = torch.arange(self.node_count, device=rows.device)
                rows, cols = torch.cat([eye, rows]), torch.cat([eye, cols])
            vals = torch.ones(rows.numel(), dtype=torch.float32, device=rows.device)
            self._adjacency = torch.sparse_coo_tensor(torch.stack([rows, cols]), vals,
                                                      (self.node_count, self.node_count)).coalesce()
        return self._adjacency
