"""Mirror of the `GraphDataset` attributes the hot-path layers consume
(/root/reference/Dataset.py:141-186 and the lazy `.graph` / `.hypergraph` properties :78-96).

The reference parses CSV search logs in Python; this container is filled either from a
`synth.SearchLogSet` (arrays) or from the reference's on-disk files, and builds its
hypergraph on the device through `ihgnn_b200.graph.PpsHyperGraph`.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from .graph import Pps2DGraph, PpsHyperGraph
from .synth import SearchLogSet


class GraphDataset:
    """Holds counts, index tensors and the lazily built hypergraph, attribute-compatible with
    the reference's GraphDataset for everything RawGnn and the layers read."""

    device: torch.device

    def __init__(self, user_count: int, query_count: int, item_count: int, vocab_size: int,
                 bag_words: np.ndarray, bag_offsets: np.ndarray,
                 pos_user: np.ndarray, pos_query: np.ndarray, pos_item: np.ndarray,
                 device, graph_type=PpsHyperGraph):
        device = torch.device(device)
        GraphDataset.device = device                       # Dataset.py:135
        self.graph_type = graph_type
        self.user_count, self.query_count = int(user_count), int(query_count)
        self.item_count, self.vocab_size = int(item_count), int(vocab_size)
        self.node_count = self.user_count + self.query_count + self.item_count   # Dataset.py:145
        self.query_start_index_in_graph = self.user_count                         # :146
        self.item_start_index_in_graph = self.user_count + self.query_count       # :147
        # one-hot value == index + 1 (row 0 of every table is padding)   Dataset.py:149-155
        self.users_onehot = torch.arange(1, 1 + self.user_count, device=device)
        self.items_onehot = torch.arange(1, 1 + self.item_count, device=device)
        self.vocabulary_onehot = torch.arange(1, 1 + self.vocab_size, device=device)
        # EmbeddingBag inputs: flat word ids (+1) and per-query start offsets   Dataset.py:165-186
        self.queries_for_embeddingbag = torch.as_tensor(np.asarray(bag_words, dtype=np.int64), device=device)
        self.queries_offset_for_embeddingbag = torch.as_tensor(np.asarray(bag_offsets, dtype=np.int64), device=device)
        self.queries_multihot = None       # legacy sparse view of the reference; never read
        self.pos_user = np.asarray(pos_user, dtype=np.int64)
        self.pos_query = np.asarray(pos_query, dtype=np.int64)
        self.pos_item = np.asarray(pos_item, dtype=np.int64)
        self.raw_logs = None               # (log_user, log_query, log_ptr, log_items, log_flags) when known
        self._hgraph: Optional[PpsHyperGraph] = None
        self._cgraph: Optional[PpsHyperGraph] = None
        self._graph2d: Optional[Pps2DGraph] = None

    def __len__(self) -> int:
        return int(self.pos_user.shape[0])

    @property
    def hypergraph(self) -> PpsHyperGraph:                  # Dataset.py:91-96
        if self._hgraph is None:
            self._hgraph = PpsHyperGraph.from_tensors(
                self.pos_user, self.pos_query, self.pos_item,
                self.user_count, self.query_count, self.item_count, GraphDataset.device)
        return self._hgraph

    @property
    def graph2d(self) -> Pps2DGraph:                        # Dataset.py:83-90 (no self connection)
        if self._graph2d is None:
            self._graph2d = Pps2DGraph.from_hypergraph(self._compute_graph(), False)
        return self._graph2d

    @property
    def graph(self):                                        # Dataset.py:79-82
        if self.graph_type is Pps2DGraph:
            return self.graph2d
        return self._compute_graph()

    def _compute_graph(self) -> PpsHyperGraph:
        """The graph the layers convolve over.  Hyperedge ids never leave the layers (their outputs
        are node features), so this copy numbers the hyperedges by ascending user (stable): the
        user third of every edge -> node reduction then reads hyperedge rows sequentially and the
        u-row gathers of consecutive hyperedges hit the same row.  Same hypergraph, same results up
        to fp32 summation order inside query / item rows.  `hypergraph` keeps the reference's
        interaction order (bit-exact I3 / Adjacency); IHG_EDGE_ORDER=file uses it here too."""
        if self._cgraph is None:
            if os.environ.get("IHG_EDGE_ORDER", "user") == "file":
                self._cgraph = self.hypergraph
            else:
                order = np.argsort(self.pos_user, kind="stable")
                self._cgraph = PpsHyperGraph.from_tensors(
                    self.pos_user[order], self.pos_query[order], self.pos_item[order],
                    self.user_count, self.query_count, self.item_count, GraphDataset.device)
        return self._cgraph

    @classmethod
    def from_search_log(cls, log: SearchLogSet, device) -> "GraphDataset":
        words, offsets = log.bag_inputs()
        ds = cls(log.user_count, log.query_count, log.item_count, log.vocab_size, words, offsets,
                 log.pos_user, log.pos_query, log.pos_item, device)
        if getattr(log, "log_ptr", None) is not None:
            ds.raw_logs = (log.log_user, log.log_query, log.log_ptr, log.log_items, log.log_flags)
        return ds

    @classmethod
    def from_files(cls, directory: str, device, train_file: str = "train_data.csv") -> "GraphDataset":
        """Read the reference's on-disk format (graph_info.txt, queries_multihot.txt,
        train_data.csv; Dataset.py:141-200, Helpers/SearchLog.py:63-71)."""
        d = read_reference_files(directory, train_file)
        ds = cls(d["user_count"], d["query_count"], d["item_count"], d["vocab_size"], d["bag_words"],
                 d["bag_offsets"], d["pos_user"], d["pos_query"], d["pos_item"], device)
        ds.raw_logs = (d["log_user"], d["log_query"], d["log_ptr"], d["log_items"], d["log_flags"])
        return ds


def _split_ints(fields) -> np.ndarray:
    """All space-separated integers of a sequence of strings, concatenated (one C-level split)."""
    text = " ".join(fields)
    return np.array(text.split(), dtype=np.int64) if text.strip() else np.zeros(0, dtype=np.int64)


def read_search_logs(path: str):
    """One search-log CSV of the reference (header, then `user,query,search_time,items,pages,positions,
    interactions,times` with space-separated list fields; Helpers/SearchLog.py:63-75,
    SearchLogCollection.py:26-32), read column-wise instead of one Python object per row:
    (log_user [n], log_query [n], log_ptr [n+1], items [nnz], flags [nnz])."""
    users, queries, items_f, flags_f = [], [], [], []
    with open(path, encoding="utf-8") as f:
        f.readline()                                               # header (SearchLogCollection.py:29)
        for line in f:
            if not line.strip():
                continue
            parts = line.rstrip("\n").split(",")
            users.append(parts[0]); queries.append(parts[1]); items_f.append(parts[3]); flags_f.append(parts[6])
    n = len(users)
    log_user = np.array(users, dtype=np.int64) if n else np.zeros(0, np.int64)
    log_query = np.array(queries, dtype=np.int64) if n else np.zeros(0, np.int64)
    counts = np.fromiter((len(s.split()) for s in items_f), dtype=np.int64, count=n)
    log_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=log_ptr[1:])
    items, flags = _split_ints(items_f), _split_ints(flags_f)
    if items.shape[0] != log_ptr[-1] or flags.shape[0] != log_ptr[-1]:
        raise ValueError(f"{path}: items / interactions lists of unequal length")
    return log_user, log_query, log_ptr, items, flags


def read_reference_files(directory: str, train_file: str = "train_data.csv") -> dict:
    """Host-side reader of a reference dataset directory (Dataset.py:141-212): counts from
    graph_info.txt, EmbeddingBag inputs from queries_multihot.txt (word ids + 1, :165-176), the
    positive (user, query, item) interactions of `train_file` in file order (flag > 0, one per
    interacted item, :196-200 / SearchLog.py:199-207), plus the raw logs for negative-item lookups."""
    with open(os.path.join(directory, "graph_info.txt"), encoding="utf-8") as f:
        U, Q, I, V = (int(p) for p in f.readline().split())
    with open(os.path.join(directory, "queries_multihot.txt"), encoding="utf-8") as f:
        lines = [ln.strip() for ln in f]
    while lines and len(lines) > Q and not lines[-1]:
        lines.pop()                                                # trailing blank lines are not queries
    lens = np.fromiter((len(ln.split()) for ln in lines), dtype=np.int64, count=len(lines))
    offsets = np.zeros(len(lines), dtype=np.int64)
    if len(lines) > 1:
        np.cumsum(lens[:-1], out=offsets[1:])
    words = _split_ints(lines) + 1                                 # Dataset.py:169: index + 1, row 0 is padding
    log_user, log_query, log_ptr, items, flags = read_search_logs(os.path.join(directory, train_file))
    per_log = np.diff(log_ptr)
    keep = flags > 0
    return {"user_count": U, "query_count": Q, "item_count": I, "vocab_size": V,
            "bag_words": words, "bag_offsets": offsets,
            "pos_user": np.repeat(log_user, per_log)[keep], "pos_query": np.repeat(log_query, per_log)[keep],
            "pos_item": items[keep],
            "log_user": log_user, "log_query": log_query, "log_ptr": log_ptr, "log_items": items, "log_flags": flags}


def logged_negative_lists(log_user, log_query, log_ptr, log_items, log_flags, pos_user, pos_query,
                          query_count: int):
    """`neg_items_for_user_query_pair` (Dataset.py:196-209) as a CSR: for every (user, query) pair that
    occurs in the logs, the items it was shown without interacting (flag <= 0), in log order, duplicates
    kept.  Returns (pos_pair [E] = pair id of every positive, neg_ptr [P+1], neg_items)."""
    log_key = np.asarray(log_user, dtype=np.int64) * int(query_count) + np.asarray(log_query, dtype=np.int64)
    keys, log_pair = np.unique(log_key, return_inverse=True)
    per_log = np.diff(np.asarray(log_ptr, dtype=np.int64))
    inc_pair = np.repeat(log_pair, per_log)
    neg = np.asarray(log_flags) <= 0
    order = np.argsort(inc_pair[neg], kind="stable")
    neg_items = np.asarray(log_items, dtype=np.int64)[neg][order]
    neg_ptr = np.zeros(keys.shape[0] + 1, dtype=np.int64)
    np.cumsum(np.bincount(inc_pair[neg], minlength=keys.shape[0]), out=neg_ptr[1:])
    pos_key = np.asarray(pos_user, dtype=np.int64) * int(query_count) + np.asarray(pos_query, dtype=np.int64)
    pos_pair = np.searchsorted(keys, pos_key)
    if pos_pair.size and not np.array_equal(keys[np.minimum(pos_pair, keys.shape[0] - 1)], pos_key):
        raise ValueError("a positive interaction's (user, query) pair does not occur in the logs")
    return pos_pair.astype(np.int64), neg_ptr, neg_items


def read_test_logs(path: str):
    """`TestSearchLogDataLoader.logs` (Dataset.py:301-318): (user, query, interacted items, None, True)
    for every search log with at least one interaction."""
    log_user, log_query, log_ptr, items, flags = read_search_logs(path)
    logs = []
    for k in range(log_user.shape[0]):
        a, b = int(log_ptr[k]), int(log_ptr[k + 1])
        hit = items[a:b][flags[a:b] > 0]
        if hit.size:
            logs.append((int(log_user[k]), int(log_query[k]), [int(x) for x in hit], None, True))
    return logs


class DeviceBatchSampler:
    """Device-side replacement of `DataLoader(dataset, batch_size, shuffle=True,
    collate_fn=GraphDataset.collate_fn)` (/root/reference/Main.py:152 with Dataset.py:107-119,
    :260-293): iterating yields the 8-tuple `collate_fn` returns -- (users, queries, items, flags,
    neg_users, neg_queries, neg_items, neg_flags), all int64 on the device -- one kernel launch per
    batch (`ihg_sample_batch`), no Python per-positive loop and no host->device copies.  One pass =
    one epoch over a fresh permutation of the positives (the last batch may be short, as with
    drop_last=False).  Seeded, hence reproducible; the reference is unseeded, so parity with it is
    distributional (uniform, distinct negatives per positive; the positive item is not excluded)."""

    def __init__(self, dataset: GraphDataset, batch_size: int = 100, neg_sample_size: int = 10, seed: int = 0,
                 nonrandom_neg_sample_size: int = 0):
        """`neg_sample_size` random negatives per positive (Gs.random_negative_sample_size) plus
        `nonrandom_neg_sample_size` taken from the items the (user, query) pair was shown without
        interacting (Gs.non_random_negative_sample_size, 0 in the reference's defaults; Dataset.py:110-119);
        the latter needs the raw logs (`dataset.raw_logs`: `from_files`, or a synthetic log with negatives)."""
        from . import _lib
        self._lib = _lib
        dev = GraphDataset.device
        if torch.device(dev).type != "cuda":
            raise RuntimeError("DeviceBatchSampler samples on the GPU; there is no CPU fallback")
        self.dataset, self.batch_size = dataset, int(batch_size)
        self.nonrand = int(nonrandom_neg_sample_size)
        self.neg = int(neg_sample_size) + self.nonrand                      # Dataset.py:139
        self.pos_pair = self.neg_ptr = self.neg_items = None
        if self.nonrand > 0:
            if dataset.raw_logs is None:
                raise ValueError("nonrandom_neg_sample_size > 0 needs dataset.raw_logs (the logged non-interactions)")
            pp, nptr, nit = logged_negative_lists(*dataset.raw_logs, dataset.pos_user, dataset.pos_query,
                                                  dataset.query_count)
            self.pos_pair = torch.as_tensor(pp, dtype=torch.int64, device=dev)
            self.neg_ptr = torch.as_tensor(nptr, dtype=torch.int64, device=dev)
            self.neg_items = torch.as_tensor(nit if nit.size else np.zeros(1, np.int64), dtype=torch.int64, device=dev)
        self.seed, self.step, self.device = int(seed), 0, dev
        self.pos_user = torch.as_tensor(dataset.pos_user, dtype=torch.int64, device=dev)
        self.pos_query = torch.as_tensor(dataset.pos_query, dtype=torch.int64, device=dev)
        self.pos_item = torch.as_tensor(dataset.pos_item, dtype=torch.int64, device=dev)
        self._gen = torch.Generator(device=dev)
        self._gen.manual_seed(self.seed)

    def __len__(self) -> int:
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def sample(self, pick: torch.Tensor):
        """The 8-tuple for the positives `pick` (int64 indices into the interaction list)."""
        L = self._lib
        L.require_cuda(pick)
        pick = pick.to(torch.int64).contiguous()
        B, K = int(pick.numel()), self.neg
        mk = lambda n: torch.empty(n, dtype=torch.int64, device=self.device)
        out = [mk(B) for _ in range(4)] + [mk(B * K) for _ in range(4)]
        L.call("ihg_sample_batch", L.ptr(self.pos_user), L.ptr(self.pos_query), L.ptr(self.pos_item), L.ptr(pick),
               B, K, self.dataset.item_count, self.seed & (2 ** 64 - 1), self.step, *(L.ptr(t) for t in out),
               L.ptr(self.pos_pair), L.ptr(self.neg_ptr), L.ptr(self.neg_items), self.nonrand, L.stream_ptr())
        self.step += 1
        return tuple(out)

    def __iter__(self):
        perm = torch.randperm(len(self.dataset), generator=self._gen, device=self.device)
        for s in range(0, int(perm.numel()), self.batch_size):
            yield self.sample(perm[s:s + self.batch_size])
