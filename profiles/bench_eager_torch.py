"""Same-box bar: the eager-PyTorch form of the conv stack -- the oracle's torch ops
(index gathers, cat, Linear, torch.sparse.mm: the ops the reference itself runs) -- executed ON
THE GPU, timed with CUDA events, next to this library on the same workload.  Baseline only; the oracle is
test infrastructure and is never on the product path.

    python profiles/bench_eager_torch.py [--workload amazon-full] [--steps 5]
"""
import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from ihgnn_b200 import synth  # noqa: E402
from oracle import ihgnn_oracle as orc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="amazon-full")
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    w = synth.WORKLOADS[a.workload]
    layers, d = w["layers"], w["dim"]
    log = synth.make_workload(a.workload)
    dev = torch.device("cuda", 0)
    g = orc.build_hypergraph(log.pos_user, log.pos_query, log.pos_item, log.user_count, log.query_count, log.item_count)
    gen = torch.Generator().manual_seed(0)
    state = {}
    for k in range(layers):
        K = 7 if k == 0 else 3
        state[f"gnn_{k}.feature_interactor.aggregation.weight"] = (torch.rand(d, K * d, generator=gen) - 0.5) * (2 / (K * d) ** 0.5)
        state[f"gnn_{k}.feature_interactor.aggregation.bias"] = (torch.rand(d, generator=gen) - 0.5) * 0.1
        state[f"gnn_{k}.feature_transform.weight"] = (torch.rand(d, d, generator=gen) - 0.5) * (2 / d ** 0.5)
        state[f"gnn_{k}.feature_transform.bias"] = (torch.rand(d, generator=gen) - 0.5) * 0.1
    words, offsets = log.bag_inputs()
    m = orc.OracleModel(state, g, torch.from_numpy(words), torch.from_numpy(offsets), log.user_count,
                        log.query_count, log.item_count, layer_type="IHGNN", layer_count=layers, order=3)
    # move the oracle's tensors to the GPU: the same eager ops now run as ATen / cuSPARSE kernels
    m.params = {k: v.detach().to(dev).requires_grad_(True) for k, v in m.params.items()}
    m.adjacency = m.adjacency.to(dev)
    m.dv_neg_1 = m.dv_neg_1.to(dev)
    g.I3 = g.I3.to(dev)
    x = (torch.randn(log.node_count, d, generator=gen) * 0.05).to(dev)
    for _ in range(2):
        orc.conv_fwd_bwd(m, x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        orc.conv_fwd_bwd(m, x)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / a.steps
    print(json.dumps({"what": "eager PyTorch ops of the reference path on the same GPU (oracle on cuda:0)",
                      "workload": a.workload, "ms_per_step": t * 1e3,
                      "hyperedge_layers_per_s": log.edge_count * layers / t,
                      "peak_memory_GB": torch.cuda.max_memory_allocated() / 1e9}))


if __name__ == "__main__":
    main()
