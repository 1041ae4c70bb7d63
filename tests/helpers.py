"""Shared helpers for the test-suite (tests only; may import oracle/)."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from oracle import ihgnn_oracle as orc  # noqa: E402

# north_star tolerance: fp32 outputs and gradients within 1e-5 max-norm relative of the
# reference PyTorch path (elementwise relative error is meaningless near
# the exact zeros of isolated nodes, hence max-norm).
REL_TOL = 1e-5


def max_rel(a, b) -> float:
    """max|a-b| / max|b| (max-norm relative error against reference b)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    denom = np.abs(b).max() if b.size else 0.0
    if denom == 0.0:
        return float(np.abs(a).max()) if a.size else 0.0
    return float(np.abs(a - b).max() / denom)


def state_of(golden) -> dict:
    return {k[len("state."):]: torch.from_numpy(golden[k]) for k in golden if k.startswith("state.")}


def oracle_graph(golden):
    U, Q, I, V, E = (int(x) for x in golden["counts"])
    return orc.build_hypergraph(golden["pos_user"], golden["pos_query"], golden["pos_item"], U, Q, I)


def oracle_model(golden, dtype=torch.float32) -> "orc.OracleModel":
    U, Q, I, V, E = (int(x) for x in golden["counts"])
    g = oracle_graph(golden)
    return orc.OracleModel(state_of(golden), g,
                           torch.from_numpy(golden["bag_words"]), torch.from_numpy(golden["bag_offsets"]),
                           U, Q, I, layer_type=str(golden["cfg.gnn"]), layer_count=int(golden["cfg.L"]),
                           order=int(golden["cfg.order"]), lambda_muq=float(golden["cfg.lambda_muq"]),
                           dtype=dtype, cosine=bool(golden.get("cfg.cosine", False)),
                           query_activation=(str(golden["cfg.query_activation"]) or None)
                           if "cfg.query_activation" in golden else None)


def batch_of(golden):
    return (torch.from_numpy(golden["batch.users"]), torch.from_numpy(golden["batch.queries"]),
            torch.from_numpy(golden["batch.items"]), torch.from_numpy(golden["batch.flags"]))
