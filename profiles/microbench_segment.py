"""Dev tool: time ihg_segment_reduce (and reference points) on a workload's hypergraph.
    python profiles/microbench_segment.py amazon-full 64 [chunk_len]
Variants are selected with IHG_SEG_VARIANT (see csrc/segment_reduce.cu)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ihgnn_b200 import synth, functional as F_
from ihgnn_b200.graph import PpsHyperGraph

name, d = sys.argv[1], int(sys.argv[2])
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 128
log = synth.make_workload(name)
g = PpsHyperGraph.from_tensors(log.pos_user, log.pos_query, log.pos_item, log.user_count, log.query_count,
                               log.item_count, "cuda:0", chunk_len=chunk)
E, N = g.EdgeCount, g.node_count
ef = torch.randn(E, d, device="cuda:0")
x = torch.randn(N, d, device="cuda:0")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


nbytes = 3 * E * (4 + 4 * d) + N * (4 * d + 16)
t = timeit(lambda: F_.segment_reduce(g.plan, ef, d, row_scale=g.dv_inv))
print(f"{name} d={d} chunk={chunk} variant={os.environ.get('IHG_SEG_VARIANT','0')}: segment_reduce {t*1e3:.1f} us  "
      f"{nbytes/t/1e6:.0f} GB/s  (n_seg={g.plan.n_seg} n_split={g.plan.n_split} n_part={g.plan.n_part})")
if os.environ.get("IHG_GS_ONLY"):
    t = timeit(lambda: F_.edge_gather_sum(x, g.i3))
    print(f"{name} d={d} gs_variant={os.environ.get('IHG_GS_VARIANT','0')}: edge_gather_sum {t*1e3:.1f} us  {E*(12+16*d)/t/1e6:.0f} GB/s algorithmic")
    t = timeit(lambda: F_.edge_gather_sum(x, g.i3, node_scale=g.dv_inv))
    print(f"    with node_scale: {t*1e3:.1f} us")
    sys.exit(0)
if os.environ.get("IHG_SEG_REF"):
    col = g.col.to(torch.int64)
    t = timeit(lambda: torch.index_select(ef, 0, col))
    print(f"   torch.index_select(ef, col) {t*1e3:.1f} us  read {3*E*4*d/t/1e6:.0f} GB/s (+ same written)")
    t = timeit(lambda: F_.edge_gather_sum(x, g.i3))
    print(f"   edge_gather_sum {t*1e3:.1f} us  {E*(12+16*d)/t/1e6:.0f} GB/s algorithmic")
    big = torch.empty(1 << 28, device="cuda:0"); big2 = torch.empty_like(big)
    t = timeit(lambda: big2.copy_(big))
    print(f"   copy 1 GiB {t*1e3:.1f} us  {2*big.numel()*4/t/1e6:.0f} GB/s")
