"""Dev tool: CUDA-time breakdown of the end-to-end training step (torch.profiler), amazon-full."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ihgnn_b200 import synth, GraphDataset, RawGnn, IHGNNLayer, HemPredictionLayer
name = sys.argv[1] if len(sys.argv) > 1 else "amazon-full"
cfg = synth.WORKLOADS[name]
log = synth.make_workload(name)
dev = "cuda:0"
ds = GraphDataset.from_search_log(log, dev)
model = RawGnn(device=torch.device(dev), dataset=ds, embedding_size=cfg["dim"], gnn_layer_type=IHGNNLayer,
               gnn_layer_count=cfg["layers"], feature_interaction_order=3, phase2_attention=False,
               predictions=HemPredictionLayer, lambda_muq=0.5).to(dev)
opt = torch.optim.Adam(model.parameters(), 1e-3, fused=True)
rng = np.random.default_rng(0)
B, NEG = 100, 10
pick = rng.integers(0, log.edge_count, size=B)
users = torch.from_numpy(np.concatenate([log.pos_user[pick], np.repeat(log.pos_user[pick], NEG)])).to(dev)
queries = torch.from_numpy(np.concatenate([log.pos_query[pick], np.repeat(log.pos_query[pick], NEG)])).to(dev)
items = torch.from_numpy(np.concatenate([log.pos_item[pick], rng.integers(0, log.item_count, size=B * NEG)])).to(dev)
flags = torch.cat([torch.ones(B), torch.zeros(B * NEG)]).to(dev)
def step():
    loss = torch.nn.functional.binary_cross_entropy_with_logits(model(users, queries, items), flags)
    opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 5, e.count / 5) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"{name}: total CUDA kernel time per step {tot/1e3:.3f} ms")
for k, t, c in rows[:40]:
    print(f"{t:9.1f} us {c:5.1f}x  {k[:110]}")
