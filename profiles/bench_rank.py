"""Inference-ranking microbenchmark (BASELINE.json configs[4] shape on ONE GPU's share):
score C candidate items per (user, query) and keep the top-k, `ihg_rank_topk`.

    python profiles/bench_rank.py [--queries 131072] [--cand 1000] [--dim 192] [--items 60000]

Prints one JSON line: queries/s, candidate rows/s and the algorithmic GB/s
(8 D + C (8 + 4 D + 4) + 12 k bytes per query) against MEASURED_PEAKS.json; also the all-items mode
(the reference's evaluation loop, Dataset.py:324-329 + Metrics.py:60-61) at a smaller batch.
"""
import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from ihgnn_b200 import functional as F_  # noqa: E402


def timed(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=131072)
    ap.add_argument("--cand", type=int, default=1000)
    ap.add_argument("--dim", type=int, default=192)
    ap.add_argument("--items", type=int, default=60000)
    ap.add_argument("--users", type=int, default=200000)
    ap.add_argument("--qcount", type=int, default=1000)
    ap.add_argument("--k", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    U, Q, I, D, B, C, k = a.users, a.qcount, a.items, a.dim, a.queries, a.cand, a.k
    feat = torch.randn(U + Q + I, D, device=dev, generator=g)
    bias = torch.randn(I, device=dev, generator=g)
    users = torch.randint(0, U, (B,), device=dev, generator=g)
    queries = torch.randint(0, Q, (B,), device=dev, generator=g)
    cand = torch.randint(0, I, (B, C), device=dev, generator=g)
    kw = dict(query_row0=U, item_row0=U + Q, item_count=I, k=k)
    t = timed(lambda: F_.rank_topk(feat, users, queries, bias, 0.5, candidates=cand, **kw))
    nbytes = B * (8 * D + C * (8 + 4 * D + 4) + 12 * k)
    peak = 6650.0
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"])
    except OSError:
        pass
    Ba = 2048
    ta = timed(lambda: F_.rank_topk(feat, users[:Ba], queries[:Ba], bias, 0.5, candidates=None, **kw), reps=3)
    # the reference's loop on the same device: one forward + full sort + top-10 copy per search
    fi = feat[U + Q:]

    def ref_loop(n=64):
        for b in range(n):
            m = 0.5 * feat[queries[b] + U] + 0.5 * feat[users[b]]
            out = (fi * m).sum(1) + bias
            _, idx = torch.sort(out, descending=True)
            idx[:10].cpu()
    tr = timed(ref_loop, warm=1, reps=3) / 64
    print(json.dumps({
        "kernel": "rank_topk", "queries": B, "candidates": C, "dim": D, "items": I, "k": k,
        "ms": t * 1e3, "queries_per_s": B / t, "candidate_rows_per_s": B * C / t,
        "algorithmic_GBps": nbytes / t / 1e9, "frac_of_hbm_peak": nbytes / t / 1e9 / peak,
        "note": "item table %.0f MB is L2-resident: the gathers are served by L2, not HBM" % (I * D * 4 / 1e6),
        "all_items_mode": {"queries": Ba, "ms": ta * 1e3, "queries_per_s": Ba / ta,
                           "item_rows_per_s": Ba * I / ta, "GBps": Ba * I * (4 * D + 4) / ta / 1e9},
        "eager_torch_loop_same_gpu": {"ms_per_query": tr * 1e3, "queries_per_s": 1 / tr,
                                      "what": "per search: (F_items * m).sum(1) + bias, torch.sort, idx[:10].cpu()"},
    }))


if __name__ == "__main__":
    main()
