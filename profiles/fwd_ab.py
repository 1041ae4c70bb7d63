import os, sys, torch
sys.path.insert(0, "/root/repo")
from ihgnn_b200 import synth
from ihgnn_b200.dataset import GraphDataset
from ihgnn_b200.layers import FeatureInteractor
from ihgnn_b200 import _lib
name = sys.argv[1]
log = synth.make_workload(name); d = synth.WORKLOADS[name]["dim"]
ds = GraphDataset.from_search_log(log, "cuda:0")
fi = FeatureInteractor(ds, 3, d, d).to("cuda:0")
x = torch.randn(ds.node_count, d, device="cuda:0")
prof = _lib.KernelProfiler() if hasattr(_lib, "KernelProfiler") else None
with torch.no_grad():
    for _ in range(5): fi(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fi(x); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
ts.sort()
print(f"{name} prewait={os.environ.get('IHG_TS_PREWAIT','0')} ss={os.environ.get('IHG_FWD_SS','-')}: FeatureInteractor fwd median {ts[10]*1e3:.0f} us min {ts[0]*1e3:.0f} us")
