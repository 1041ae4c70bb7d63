"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports
every symbol include/ihgnn_b200.h declares (no compute calls -- there is no GPU here), and the
product refuses to run on CPU tensors."""
import os
import re

import pytest
import torch

from conftest import REPO


@pytest.fixture(scope="module")
def lib():
    from ihgnn_b200 import build, _lib
    build.build()
    return _lib.lib()


def _header_symbols():
    with open(os.path.join(REPO, "include", "ihgnn_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ihg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    from ihgnn_b200 import _lib
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ihgnn_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in ihgnn_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_error_channel(lib):
    assert lib.ihg_abi_version() == 4
    assert lib.ihg_last_error() is not None
    # argument validation happens before any CUDA call, so it can be exercised without a GPU
    rc = lib.ihg_node_linear(None, 0, None, 1, 64, 64, 0, None, None, 0, 10, 0, 0, None, 0, None)
    assert rc == 1 and b"null pointer" in lib.ihg_last_error()
    rc = lib.ihg_edge_interact_fwd(1, 64, 1, 64, 1, 64, 5, 1, 10, 1, 64, 64, None, 0, None)
    assert rc == 1 and b"order" in lib.ihg_last_error()


def test_new_entry_points_validate_arguments(lib):
    """ihg_rank_topk / ihg_sample_batch / ihg_two_hop_reduce / ihg_segment_reduce flags: the argument checks
    run before any CUDA call."""
    rc = lib.ihg_rank_topk(1, 64, None, 1, 4, 0, None, 10, 0, 10, 1, 0.5, 64, 33, 0, 1, 1, None)
    assert rc == 1 and b"k=33" in lib.ihg_last_error()
    rc = lib.ihg_rank_topk(1, 64, None, 1, 4, 0, None, 10, 0, 10, 1, 0.5, 62, 10, 0, 1, 1, None)
    assert rc == 1 and b"dim=62" in lib.ihg_last_error()
    rc = lib.ihg_rank_topk(None, 64, None, 1, 4, 0, None, 10, 0, 10, 1, 0.5, 64, 10, 0, 1, 1, None)
    assert rc == 1 and b"null pointer" in lib.ihg_last_error()
    rc = lib.ihg_sample_batch(1, 1, 1, 1, 8, 65, 100, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, None, None, None, 0, None)
    assert rc == 1 and b"neg_per_positive=65" in lib.ihg_last_error()
    rc = lib.ihg_sample_batch(1, 1, 1, 1, 8, 10, 5, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, None, None, None, 0, None)
    assert rc == 1 and b"distinct" in lib.ihg_last_error()
    rc = lib.ihg_sample_batch(1, 1, 1, 1, 8, 10, 100, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, None, None, None, 3, None)
    assert rc == 1 and b"logged negatives" in lib.ihg_last_error()
    rc = lib.ihg_sample_batch(1, 1, 1, 1, 8, 10, 100, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 11, None)
    assert rc == 1 and b"nonrandom_per_positive=11" in lib.ihg_last_error()
    rc = lib.ihg_two_hop_reduce(None, None, None, 64, None, 1.0, 1.0, 0.0, None, None, None, 64, 64, None)
    assert rc == 1 and b"null pointer" in lib.ihg_last_error()
    from ihgnn_b200 import _lib
    import ctypes
    csr = _lib.IhgCsr(n_rows=4, nnz=0, rowptr=1, col=None, chunk_len=128, n_seg=4, n_split=0, n_part=0, seg=1,
                      split_row=None, split_ptr=None)
    # accumulate mode without an init row
    rc = lib.ihg_segment_reduce(ctypes.byref(csr), 1, 64, 1, 0, 0, None, None, 0, None, None, None, 1, 64, 64, 1, None)
    assert rc == 1 and b"accumulate" in lib.ihg_last_error()
    # routed reductions (multi-GPU): the output ranges must cover [0, n_rows), ascending, non-null where non-empty
    starts = (ctypes.c_int64 * 3)(0, 2, 3)
    bases = (ctypes.c_void_p * 2)(1, 1)
    rc = lib.ihg_segment_reduce_routed(ctypes.byref(csr), 1, 64, 1, 0, 0, None, None, starts, bases, 2, 64, 64, None)
    assert rc == 1 and b"cover rows" in lib.ihg_last_error()
    starts = (ctypes.c_int64 * 3)(0, 2, 4)
    bases = (ctypes.c_void_p * 2)(1, None)
    rc = lib.ihg_two_hop_reduce_routed(ctypes.byref(csr), 1, 1, 64, None, 1.0, 1.0, 0.0, None, starts, bases, 2, 64, 64, None)
    assert rc == 1 and b"null destination" in lib.ihg_last_error()
    rc = lib.ihg_two_hop_reduce_routed(ctypes.byref(csr), 1, 1, 64, None, 1.0, 1.0, 0.0, None, starts, bases, 17, 64, 64, None)
    assert rc == 1 and b"output ranges" in lib.ihg_last_error()


def test_workspace_queries_are_pure(lib):
    assert lib.ihg_graph_workspace_bytes(1000, 500) > 0
    assert lib.ihg_graph_workspace_bytes(5_000_000, 500_000) > 5_000_000 * 4 * 7
    assert lib.ihg_segment_plan_workspace_bytes(1000) > 0
    assert lib.ihg_edge_interact_bwd_workspace_bytes(128, 3) >= 74 * 4 * 128 * 128 * 4
    assert lib.ihg_node_linear_wgrad_workspace_bytes(3, 64, 64) > 0


def test_product_has_no_cpu_path():
    from ihgnn_b200 import functional as F_
    from ihgnn_b200.graph import PpsHyperGraph
    with pytest.raises(RuntimeError):
        PpsHyperGraph.from_tensors([0], [0], [0], 2, 2, 2, "cpu")
    with pytest.raises(RuntimeError):
        F_.gather_rows(torch.zeros(4, 8), torch.zeros(2, dtype=torch.long), 0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(REPO, "ihgnn_b200")
    for root, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                with open(os.path.join(root, f)) as fh:
                    src = fh.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_pps2dgraph_rejects_unknown_completeness(monkeypatch):
    """graph_completeness outside uqi / uq / ui / qi raises ValueError, as Helpers/Graph.py:64-65 does
    (the four values themselves are covered on the GPU: test_graph2d_every_branch_matches_reference)."""
    from ihgnn_b200 import settings
    from ihgnn_b200.graph import Pps2DGraph
    monkeypatch.setattr(settings.Gs, "graph_completeness", "uqx", raising=False)
    with pytest.raises(ValueError):
        Pps2DGraph.from_hypergraph(None, False)
