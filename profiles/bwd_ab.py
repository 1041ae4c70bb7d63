"""Dev tool: FeatureInteractor (order 3) forward + backward on a named workload; run under
ncu --metrics gpu__time_duration.sum to get the per-kernel split of the backward."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ihgnn_b200 import synth
from ihgnn_b200.dataset import GraphDataset
from ihgnn_b200.layers import FeatureInteractor
name = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
log = synth.make_workload(name); d = synth.WORKLOADS[name]["dim"]
ds = GraphDataset.from_search_log(log, "cuda:0")
fi = FeatureInteractor(ds, 3, d, d).to("cuda:0")
x = torch.randn(ds.node_count, d, device="cuda:0", requires_grad=True)
g = torch.randn(ds.graph.EdgeCount, d, device="cuda:0")
ts = []
for _ in range(reps):
    ef = fi(x)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ef.backward(g); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
print(f"{name}: FeatureInteractor backward min {min(ts)*1e3:.0f} us")
