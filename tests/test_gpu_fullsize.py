"""Parity at BASELINE.json's full sizes (configs[1] amazon-full: E = 1.2 M, d = 64; configs[2] cikm:
E = 5 M, d = 128) through size-independent properties -- the oracle cannot run these sizes in seconds:

  * index invariants of the graph build (every hyperedge exactly once per slot, rows sorted, degrees);
  * a checksum of checksums: sum_v (H ef)[v] == 3 * sum_e ef[e];
  * the edge -> node reduction and the node -> edge gather are adjoint: <H ef, y> == <ef, H^T y>;
  * linearity of the reductions;
  * independent implementations agree: two-hop pass == gather-sum + segmented reduce; the un-hoisted
    tensor-core FeatureInteractor forward == the hoisted form; the C-ABI backward == the adjoint identity
    <dOut, J v> == <J^T dOut, v> probed with the (linear) order-1 layer;
  * renumbering the hyperedges leaves every node output unchanged (up to fp32 summation order);
  * bitwise run-to-run determinism.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
TOL = 1e-5            # north_star: fp32 within 1e-5 max-norm relative


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module", params=["amazon-full", "cikm"])
def workload(request):
    from ihgnn_b200 import synth
    from ihgnn_b200.dataset import GraphDataset
    log = synth.make_workload(request.param)
    ds = GraphDataset.from_search_log(log, DEV)
    return request.param, log, ds, synth.WORKLOADS[request.param]["dim"]


def test_graph_build_invariants(workload):
    name, log, ds, d = workload
    g = ds.hypergraph                                     # reference (file) hyperedge order
    E, N, U, Q = g.EdgeCount, g.node_count, g.user_count, g.query_count
    assert E == log.edge_count and N == log.node_count
    i3 = g.i3.to(torch.int64)
    assert torch.equal(i3[:, 0].cpu(), torch.from_numpy(log.pos_user))
    assert torch.equal(i3[:, 1].cpu() - U, torch.from_numpy(log.pos_query))
    assert torch.equal(i3[:, 2].cpu() - U - Q, torch.from_numpy(log.pos_item))
    rowptr, col = g.rowptr.to(torch.int64), g.col.to(torch.int64)
    assert int(rowptr[0]) == 0 and int(rowptr[-1]) == 3 * E and bool((rowptr[1:] >= rowptr[:-1]).all())
    assert bool((torch.bincount(col, minlength=E) == 3).all())           # every hyperedge once per slot
    deg = rowptr[1:] - rowptr[:-1]
    assert torch.equal(deg, torch.bincount(i3.reshape(-1), minlength=N))
    # ascending hyperedge ids inside every row (the order coalesce() gives): col increases except at row starts
    inc = col[1:] > col[:-1]
    starts = torch.zeros(3 * E, dtype=torch.bool, device=DEV)
    starts[rowptr[1:-1][deg[1:] > 0]] = True
    assert bool((inc | starts[1:]).all())
    # the row of incidence j holds hyperedge col[j]
    rows = torch.repeat_interleave(torch.arange(N, device=DEV), deg)
    slot = (rows >= U).to(torch.int64) + (rows >= U + Q).to(torch.int64)
    assert torch.equal(i3[col, slot], rows)
    vd = g.VertexDegrees.view(-1)
    assert torch.equal(vd, torch.where(deg == 0, torch.full_like(vd, 1e-8), deg.to(torch.float32)))
    assert torch.equal(g.dv_inv, vd.pow(-1))


def test_reductions_checksum_adjoint_linearity(workload):
    from ihgnn_b200 import functional as F_
    name, log, ds, d = workload
    g = ds.graph
    E, N = g.EdgeCount, g.node_count
    gen = torch.Generator(device=DEV).manual_seed(1)
    ef = torch.randn(E, d, device=DEV, generator=gen)
    y = torch.randn(N, d, device=DEV, generator=gen)
    s = F_.segment_reduce(g.plan, ef, d)                                  # H ef
    # checksum of checksums, per feature column, in fp64
    assert rel(s.double().sum(0), 3.0 * ef.double().sum(0)) < TOL
    hty = F_.edge_gather_sum(y, g.i3)                                     # H^T y
    lhs, rhs = (s.double() * y.double()).sum(), (ef.double() * hty.double()).sum()
    assert abs(float(lhs - rhs)) / abs(float(rhs)) < TOL
    ef2 = torch.randn(E, d, device=DEV, generator=gen)
    lin = F_.segment_reduce(g.plan, 0.5 * ef - 2.0 * ef2, d)
    assert rel(lin, 0.5 * s - 2.0 * F_.segment_reduce(g.plan, ef2, d)) < TOL
    # per-slot rows [E,3,d]: reducing slot-replicated copies equals the plain reduction
    rep = ef.unsqueeze(1).expand(E, 3, d).contiguous()
    s3 = F_.segment_reduce(g.plan, rep, d, src_row_mul=3, bounds=g.type_bounds)
    assert torch.equal(s3, s)
    # isolated nodes are exact zeros; determinism
    iso = (g.rowptr[1:] - g.rowptr[:-1]) == 0
    assert float(s[iso].abs().max()) == 0.0 if bool(iso.any()) else True
    assert torch.equal(s, F_.segment_reduce(g.plan, ef, d))


def test_two_hop_equals_gather_then_reduce(workload):
    from ihgnn_b200 import functional as F_
    name, log, ds, d = workload
    g = ds.graph
    N = g.node_count
    gen = torch.Generator(device=DEV).manual_seed(2)
    x = torch.randn(N, d, device=DEV, generator=gen)
    nbr = g.plan.two_hop_nbr(g.i3, g.type_bounds)
    a = F_.two_hop_reduce(g.plan, nbr, x, node_scale=g.dv_inv_sqrt, alpha=1.0 / 3.0, row_scale=g.dv_inv)
    b = F_.segment_reduce(g.plan, F_.edge_gather_sum(x, g.i3, node_scale=g.dv_inv_sqrt, alpha=1.0 / 3.0), d,
                          row_scale=g.dv_inv)
    assert rel(a, b) < TOL
    assert torch.equal(a, F_.two_hop_reduce(g.plan, nbr, x, node_scale=g.dv_inv_sqrt, alpha=1.0 / 3.0,
                                            row_scale=g.dv_inv))
    # pairwise-graph form (GCN): own term off == H H^T x minus deg * x
    c = F_.two_hop_reduce(g.plan, nbr, x, own=(0.0, 0.0))
    deg = (g.rowptr[1:] - g.rowptr[:-1]).to(torch.float32).view(-1, 1)
    full = F_.two_hop_reduce(g.plan, nbr, x)
    assert rel(c, full - deg * x) < TOL


def test_interactor_implementations_agree_and_order1_layer_is_adjoint_consistent(workload):
    from ihgnn_b200 import _lib
    from ihgnn_b200.layers import FeatureInteractor, IHGNNLayer, _EdgeInteractFn
    name, log, ds, d = workload
    g = ds.graph
    N = g.node_count
    torch.manual_seed(3)
    gen = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(N, d, device=DEV, generator=gen) * 0.5
    fi = FeatureInteractor(ds, 3, d, d).to(DEV)
    assert fi._full_supported()
    with torch.no_grad():
        full = fi(x)                                                      # un-hoisted tensor-core forward
        p = fi._first_order(x)
        hoisted = _EdgeInteractFn.apply(x, p, fi.aggregation.weight[:, 3 * d:], g, 3)
    assert rel(full, hoisted) < TOL
    # order-1 IHGNN layer is affine in x: J = d out / d x is a fixed matrix, so for random v, w:
    # <w, layer(x + v) - layer(x)> == <J^T w, v> where J^T w comes from the C-ABI backward
    layer = IHGNNLayer(torch.device(DEV), ds, d, d, 1, False).to(DEV)
    v = torch.randn(N, d, device=DEV, generator=gen)
    w = torch.randn(N, d, device=DEV, generator=gen)
    xr = x.clone().requires_grad_(True)
    out = layer(xr)
    out.backward(w)
    with torch.no_grad():
        diff = layer(x + v) - out.detach()
    lhs, rhs = (w.double() * diff.double()).sum(), (xr.grad.double() * v.double()).sum()
    assert abs(float(lhs - rhs)) / abs(float(rhs)) < 1e-4                  # difference of two fp32 evaluations
    assert _lib.launch_count() > 0


def test_hyperedge_renumbering_leaves_node_outputs_unchanged(workload):
    """`GraphDataset.graph` numbers the hyperedges by user, `GraphDataset.hypergraph` in file order: the
    conv stack must not care (node features only; fp32 summation order differs)."""
    from ihgnn_b200.layers import IHGNNLayer
    name, log, ds, d = workload
    torch.manual_seed(4)
    layer0 = IHGNNLayer(torch.device(DEV), ds, d, d, 3, False).to(DEV)
    layer1 = IHGNNLayer(torch.device(DEV), ds, d, d, 1, False).to(DEV)
    gen = torch.Generator(device=DEV).manual_seed(4)
    x = torch.randn(ds.node_count, d, device=DEV, generator=gen) * 0.5

    class FileOrder:                                      # same dataset, hyperedges in file order
        pass
    alt = FileOrder()
    alt.__dict__.update(ds.__dict__)
    alt.graph = ds.hypergraph
    alt.hypergraph = ds.hypergraph
    a0 = IHGNNLayer(torch.device(DEV), alt, d, d, 3, False).to(DEV)
    a1 = IHGNNLayer(torch.device(DEV), alt, d, d, 1, False).to(DEV)
    a0.load_state_dict(layer0.state_dict())
    a1.load_state_dict(layer1.state_dict())
    with torch.no_grad():
        o = layer1(layer0(x))
        oa = a1(a0(x))
    assert rel(oa, o) < TOL


def test_ranking_at_config5_shape_properties():
    """BASELINE.json configs[4] per-GPU share (1 000 candidates per search, D = 192, 60 K items): sorted
    scores, ids drawn from the search's own candidate list, idempotence (re-ranking the winners returns
    them in the same order with the same scores), and agreement with torch.topk on a slice."""
    from ihgnn_b200 import functional as F_
    U, Q, I, D, B, C, k = 200_000, 1_000, 60_000, 192, 32_768, 1_000, 10
    gen = torch.Generator(device=DEV).manual_seed(5)
    feat = torch.randn(U + Q + I, D, device=DEV, generator=gen)
    bias = torch.randn(I, device=DEV, generator=gen)
    users = torch.randint(0, U, (B,), device=DEV, generator=gen)
    queries = torch.randint(0, Q, (B,), device=DEV, generator=gen)
    cand = torch.randint(0, I, (B, C), device=DEV, generator=gen)
    kw = dict(query_row0=U, item_row0=U + Q, item_count=I)
    ids, vals = F_.rank_topk(feat, users, queries, bias, 0.5, candidates=cand, k=k, **kw)
    assert bool((vals[:, 1:] <= vals[:, :-1]).all())                                   # sortedness
    assert bool((ids.unsqueeze(2) == cand.unsqueeze(1)).any(2).all())                  # members of the own list
    ids2, vals2 = F_.rank_topk(feat, users, queries, bias, 0.5, candidates=ids, k=k, **kw)
    # idempotence (duplicate candidate ids may swap places among equal scores only)
    assert torch.equal(vals2, vals) and bool(((ids2 == ids) | (vals2 == vals)).all())
    n = 1024
    m = 0.5 * feat[queries[:n] + U] + 0.5 * feat[users[:n]]
    sc = torch.einsum("bcd,bd->bc", feat[cand[:n] + U + Q].double(), m.double()) + bias[cand[:n]].double()
    tv, ti = torch.topk(sc, k, dim=1)
    assert rel(vals[:n], tv) < TOL
    agree = (torch.gather(cand[:n], 1, ti) == ids[:n]).float().mean()
    assert float(agree) > 0.999                            # fp32 vs fp64 near-ties may swap neighbours
    again, _ = F_.rank_topk(feat, users, queries, bias, 0.5, candidates=cand, k=k, **kw)
    assert torch.equal(again, ids)
