"""Seeded synthetic search logs of Amazon / CIKM-Cup-2016 shape (numpy only, no GPU).

The generator emits (a) the positive (user, query, item) triples in *file order* -- the
order `GraphDataset.__init__` collects `pos_interactions` in
(/root/reference/Dataset.py:192-200, Helpers/SearchLog.py:199-207), which defines the
hyperedge numbering of `PpsHyperGraph.from_interactions` (Helpers/Graph.py:94-134) -- and
(b), for small cases, the reference's on-disk files so the unmodified reference consumes
the very same data:

  graph_info.txt         "U Q I V"                         Dataset.py:143-144
  queries_multihot.txt   Q lines of 0-based word ids       Dataset.py:165-176
  {train,valid,test}_data.csv   header + one search log per line, list fields
                         space-separated                   Helpers/SearchLog.py:63-75

Amazon-shaped logs carry exactly one (positive) item (PreProcess/Step1-Amazon.py:126-134);
CIKM-shaped logs carry a 5..40 item result list of which 1..3 are clicked
(PreProcess/Step1-CikmCup2016Track2.py:124-136).  Node popularity is Zipf(s) per node
type, which leaves a tail of degree-0 nodes (the 1e-8 degree path, Graph.py:120).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np


@dataclass
class SearchLogSet:
    """A synthetic search log in array form."""
    user_count: int
    query_count: int
    item_count: int
    vocab_size: int
    # query text: 0-based word ids, CSR over queries
    query_words: np.ndarray            # int64 [sum len]
    query_word_ptr: np.ndarray         # int64 [Q+1]
    # positive interactions in file order == hyperedge order
    pos_user: np.ndarray               # int64 [E]
    pos_query: np.ndarray              # int64 [E]
    pos_item: np.ndarray               # int64 [E]
    # optional full logs (small cases only): CSR over logs
    log_user: Optional[np.ndarray] = None
    log_query: Optional[np.ndarray] = None
    log_ptr: Optional[np.ndarray] = None
    log_items: Optional[np.ndarray] = None
    log_flags: Optional[np.ndarray] = None
    shape: str = "amazon"
    seed: int = 0
    extra: Dict = field(default_factory=dict)

    @property
    def node_count(self) -> int:
        return self.user_count + self.query_count + self.item_count

    @property
    def edge_count(self) -> int:
        return int(self.pos_user.shape[0])

    def bag_inputs(self):
        """EmbeddingBag inputs exactly as Dataset.py:165-186 builds them:
        flat word ids **+1** and the Q start offsets (no trailing end offset)."""
        return self.query_words + 1, self.query_word_ptr[:-1].copy()


def _zipf_sampler(rng: np.random.Generator, n: int, s: float):
    """Returns sample(k) drawing ids in [0,n) with P(rank r) ~ (r+1)^-s under a random
    rank->id permutation (popular nodes are not the low ids)."""
    if s <= 0:
        return lambda k: rng.integers(0, n, size=k, dtype=np.int64)
    w = (np.arange(1, n + 1, dtype=np.float64)) ** (-s)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    perm = rng.permutation(n).astype(np.int64)

    def sample(k: int) -> np.ndarray:
        r = np.searchsorted(cdf, rng.random(k), side="right")
        np.minimum(r, n - 1, out=r)
        return perm[r]

    return sample


def make_search_log(user_count: int, query_count: int, item_count: int, edge_count: int,
                    vocab_size: int = 1000, shape: str = "amazon", seed: int = 0,
                    zipf: float = 0.8, with_negatives: bool = False) -> SearchLogSet:
    """Generate a synthetic log with exactly `edge_count` positive interactions.

    `with_negatives=True` also materialises the full per-log item lists (needed only to
    write the reference's CSV files); leave it off for multi-million-edge workloads.
    """
    assert shape in ("amazon", "cikm")
    rng = np.random.default_rng(seed)
    su = _zipf_sampler(rng, user_count, zipf)
    sq = _zipf_sampler(rng, query_count, zipf)
    si = _zipf_sampler(rng, item_count, zipf)

    # query text: 1..7 words per query
    qlen = rng.integers(1, 8, size=query_count, dtype=np.int64)
    qptr = np.zeros(query_count + 1, dtype=np.int64)
    np.cumsum(qlen, out=qptr[1:])
    qwords = rng.integers(0, vocab_size, size=int(qptr[-1]), dtype=np.int64)

    E = int(edge_count)
    if shape == "amazon":
        n_logs = E
        clicks = np.ones(n_logs, dtype=np.int64)
    else:
        # 1..3 clicked items per log; draw until E positives are covered, then trim
        n_logs = max(1, (E + 1) // 2 + 8)
        clicks = rng.integers(1, 4, size=n_logs, dtype=np.int64)
        while clicks.sum() < E:
            clicks = np.concatenate([clicks, rng.integers(1, 4, size=n_logs, dtype=np.int64)])
        csum = np.cumsum(clicks)
        n_logs = int(np.searchsorted(csum, E, side="left")) + 1
        clicks = clicks[:n_logs].copy()
        clicks[-1] -= int(csum[n_logs - 1] - E)
    lu = su(n_logs)
    lq = sq(n_logs)
    pos_user = np.repeat(lu, clicks)
    pos_query = np.repeat(lq, clicks)
    pos_item = si(E)
    out = SearchLogSet(user_count, query_count, item_count, vocab_size, qwords, qptr,
                       pos_user, pos_query, pos_item, shape=shape, seed=seed)
    if with_negatives:
        if shape == "amazon":
            out.log_user, out.log_query = lu, lq
            out.log_ptr = np.arange(n_logs + 1, dtype=np.int64)
            out.log_items = pos_item.copy()
            out.log_flags = np.ones(E, dtype=np.int64)
        else:
            # result list of 5..40 items; the clicked ones sit at random positions
            n_items = np.maximum(rng.integers(5, 41, size=n_logs, dtype=np.int64), clicks)
            ptr = np.zeros(n_logs + 1, dtype=np.int64)
            np.cumsum(n_items, out=ptr[1:])
            items = si(int(ptr[-1]))
            flags = np.zeros(int(ptr[-1]), dtype=np.int64)
            cptr = np.zeros(n_logs + 1, dtype=np.int64)
            np.cumsum(clicks, out=cptr[1:])
            for l in range(n_logs):
                k = int(clicks[l])
                pos = np.sort(rng.choice(int(n_items[l]), size=k, replace=False))
                flags[ptr[l] + pos] = 1
                items[ptr[l] + pos] = pos_item[cptr[l]:cptr[l] + k]
            out.log_user, out.log_query, out.log_ptr = lu, lq, ptr
            out.log_items, out.log_flags = items, flags
    return out


def _write_logs_csv(path: str, log_user, log_query, log_ptr, log_items, log_flags) -> None:
    """One line per search log in SearchLog.parse's format (Helpers/SearchLog.py:63-75)."""
    with open(path, "w", encoding="utf-8") as f:
        f.write("user,query,search_time,items,pages,positions,interactions,times\n")
        for l in range(len(log_user)):
            a, b = int(log_ptr[l]), int(log_ptr[l + 1])
            items = log_items[a:b]
            flags = log_flags[a:b]
            n = b - a
            stime = f"2016-01-01T00:{(l // 60) % 60:02d}:{l % 60:02d}"
            times = [stime if fl > 0 else "NA" for fl in flags]
            f.write(",".join([
                str(int(log_user[l])), str(int(log_query[l])), stime,
                " ".join(str(int(x)) for x in items),
                " ".join("1" for _ in range(n)),
                " ".join(str(p) for p in range(n)),
                " ".join(str(int(x)) for x in flags),
                " ".join(times),
            ]))
            f.write("\n")


def write_reference_files(log: SearchLogSet, directory: str, eval_logs: int = 8) -> None:
    """Write graph_info.txt / queries_multihot.txt / {train,valid,test}_data.csv so that
    `GraphDataset(...)` (Dataset.py:121-225) and `TestSearchLogDataLoader` (:297-322) of the
    unmodified reference read this very log."""
    assert log.log_ptr is not None, "generate with with_negatives=True to write files"
    os.makedirs(directory, exist_ok=True)
    with open(os.path.join(directory, "graph_info.txt"), "w", encoding="utf-8") as f:
        f.write(f"{log.user_count} {log.query_count} {log.item_count} {log.vocab_size}\n")
    with open(os.path.join(directory, "queries_multihot.txt"), "w", encoding="utf-8") as f:
        for q in range(log.query_count):
            a, b = int(log.query_word_ptr[q]), int(log.query_word_ptr[q + 1])
            f.write(" ".join(str(int(w)) for w in log.query_words[a:b]) + "\n")
    _write_logs_csv(os.path.join(directory, "train_data.csv"),
                    log.log_user, log.log_query, log.log_ptr, log.log_items, log.log_flags)
    # valid / test: a few single-positive logs re-drawn from the train positives
    rng = np.random.default_rng(log.seed + 7919)
    for name in ("valid_data.csv", "test_data.csv"):
        n = min(eval_logs, log.edge_count)
        pick = rng.choice(log.edge_count, size=n, replace=False)
        _write_logs_csv(os.path.join(directory, name),
                        log.pos_user[pick], log.pos_query[pick],
                        np.arange(n + 1, dtype=np.int64), log.pos_item[pick],
                        np.ones(n, dtype=np.int64))


# Workloads named by BASELINE.json `configs` (shapes per SURVEY.md section 8).
WORKLOADS: Dict[str, Dict] = {
    # configs[0]: small Amazon-shaped log, the reference's own CPU-runnable case
    "amazon-small": dict(user_count=10_000, query_count=5_000, item_count=20_000,
                         edge_count=100_000, vocab_size=8_000, shape="amazon",
                         layers=2, dim=64, seed=1),
    # configs[1]: full Amazon-subset shape on 1 B200 (the bench default)
    "amazon-full": dict(user_count=200_000, query_count=1_000, item_count=60_000,
                        edge_count=1_200_000, vocab_size=20_000, shape="amazon",
                        layers=2, dim=64, seed=2),
    # configs[2]: CIKM Cup 2016 Track2-shaped, ~5M hyperedges, 3 layers, d=128
    "cikm": dict(user_count=250_000, query_count=60_000, item_count=190_000,
                 edge_count=5_000_000, vocab_size=50_000, shape="cikm",
                 layers=3, dim=128, seed=3),
    # configs[3]: scaled synthetic hypergraph, 100 M hyperedges over 50 M table rows, multi-GPU only
    # (generated on the device: `make_workload(..., device=...)`)
    "scaled": dict(user_count=38_000_000, query_count=200_000, item_count=11_800_000,
                   edge_count=100_000_000, vocab_size=50_000, shape="amazon",
                   layers=2, dim=64, seed=4),
}


def make_search_log_on_device(user_count: int, query_count: int, item_count: int, edge_count: int, device,
                              vocab_size: int = 1000, shape: str = "amazon", seed: int = 0,
                              zipf: float = 0.8) -> SearchLogSet:
    """The same kind of log as `make_search_log` -- Zipf(s) popularity per node type under a random
    rank -> id permutation, Amazon (one positive per log) or CIKM (1..3 clicked items per log) shape --
    drawn with torch on `device` for logs too large for the host generator (BASELINE.json configs[3]:
    10^8 hyperedges take minutes in numpy, ~0.1 s on the GPU).  `pos_user / pos_query / pos_item` are int64
    torch tensors ON THE DEVICE; the (small) query texts are numpy as usual.  Seeded: every rank of a
    multi-GPU job draws the identical log on its own GPU.  A different stream than the numpy generator's,
    so the two produce different logs of the same distribution."""
    import torch
    assert shape in ("amazon", "cikm")
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))

    def sampler(n: int):
        if zipf <= 0:
            return lambda k: torch.randint(0, n, (k,), generator=gen, device=dev)
        w = torch.arange(1, n + 1, dtype=torch.float64, device=dev).pow_(-zipf)
        cdf = torch.cumsum(w, 0)
        cdf /= cdf[-1].clone()
        perm = torch.randperm(n, generator=gen, device=dev)

        def sample(k: int):
            r = torch.searchsorted(cdf, torch.rand(k, generator=gen, device=dev, dtype=torch.float64), right=True)
            return perm[r.clamp_(max=n - 1)]
        return sample

    su, sq, si = sampler(user_count), sampler(query_count), sampler(item_count)
    rng = np.random.default_rng(seed)
    qlen = rng.integers(1, 8, size=query_count, dtype=np.int64)
    qptr = np.zeros(query_count + 1, dtype=np.int64)
    np.cumsum(qlen, out=qptr[1:])
    qwords = rng.integers(0, vocab_size, size=int(qptr[-1]), dtype=np.int64)
    E = int(edge_count)
    if shape == "amazon":
        pos_user, pos_query = su(E), sq(E)
    else:
        n_logs = E                                           # >= E positives for sure; trimmed below
        clicks = torch.randint(1, 4, (n_logs,), generator=gen, device=dev)
        csum = torch.cumsum(clicks, 0)
        n_logs = int(torch.searchsorted(csum, torch.tensor([E], device=dev), right=False)) + 1
        clicks = clicks[:n_logs].clone()
        clicks[-1] -= int(csum[n_logs - 1]) - E
        lu, lq = su(n_logs), sq(n_logs)
        pos_user, pos_query = torch.repeat_interleave(lu, clicks), torch.repeat_interleave(lq, clicks)
    pos_item = si(E)
    return SearchLogSet(user_count, query_count, item_count, vocab_size, qwords, qptr,
                        pos_user, pos_query, pos_item, shape=shape, seed=seed)


def make_workload(name: str, scale: float = 1.0, with_negatives: bool = False, device=None) -> SearchLogSet:
    """Instantiate a named workload; `scale` multiplies every count (multi-GPU weak scaling
    and bounded CPU-baseline samples use it).  `device`: draw the log with torch on that device
    (`make_search_log_on_device`) instead of numpy on the host."""
    w = dict(WORKLOADS[name])
    layers, dim, seed = w.pop("layers"), w.pop("dim"), w.pop("seed")
    if scale != 1.0:
        for k in ("user_count", "query_count", "item_count", "edge_count"):
            w[k] = max(4, int(round(w[k] * scale)))
    if device is not None:
        log = make_search_log_on_device(device=device, seed=seed, **w)
    else:
        log = make_search_log(seed=seed, with_negatives=with_negatives, **w)
    log.extra.update(layers=layers, dim=dim, name=name, scale=scale)
    return log
