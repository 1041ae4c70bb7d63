"""Dev tool: in-kernel timeline of edge_interact_fwd_tc_kernel (block 0) using an instrumented
build of the library (build/libihgnn_trace.so, made from a scratch copy of csrc/ with clock64
probes; not part of the product).  Prints where producers / MMA issuer / epilogue spend time."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ihgnn_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "..", "..", "build", "libihgnn_trace.so")
from ihgnn_b200 import synth
from ihgnn_b200.dataset import GraphDataset
from ihgnn_b200.layers import FeatureInteractor

name = sys.argv[1] if len(sys.argv) > 1 else "amazon-full"
log = synth.make_workload(name)
d = synth.WORKLOADS[name]["dim"]
ds = GraphDataset.from_search_log(log, "cuda:0")
fi = FeatureInteractor(ds, 3, d, d).to("cuda:0")
x = torch.randn(ds.node_count, d, device="cuda:0")
with torch.no_grad():
    for _ in range(3):
        fi(x)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); fi(x); e.record(); torch.cuda.synchronize()
print(f"{name} d={d}: FeatureInteractor fwd {s.elapsed_time(e)*1e3:.0f} us")
lib = _lib.lib()
lib.ihg_debug_read_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = np.zeros(1 << 16, dtype=np.int64)
assert lib.ihg_debug_read_trace(buf.ctypes.data, buf.size) == 0
T = buf.reshape(8, 8192)
KC = d // 32
nblk = 7
per_tile = KC * nblk
n_it = int((T[3, :4000] > 0).sum())
tiles = n_it // per_tile
print(f"block 0: {tiles} tiles, {n_it} stages; kernel span {(T[3, :n_it].max() - T[0, 0]) / 1e3:.0f} kcycles "
      f"= {(T[3, :n_it].max() - T[0, 0]) / max(tiles,1):.0f} cycles/tile")
it = np.arange(n_it)
wait_empty = T[2, :n_it] - T[1, :n_it]
prod = T[3, :n_it] - T[2, :n_it]
first = (it % nblk) == 0
print(f"producer: wait-empty mean {wait_empty.mean():.0f} (sum/tile {wait_empty.sum()/tiles:.0f}); "
      f"produce first-stage-of-kc mean {prod[first].mean():.0f} (includes gather latency), other stages mean {prod[~first].mean():.0f}; "
      f"sum/tile {prod.sum()/tiles:.0f}")
ld_to_first = T[1, :n_it][first] - T[0, :n_it][first]
print(f"producer: load-issue -> first wait  mean {ld_to_first.mean():.0f}")
mma_wait = T[5, :n_it] - T[4, :n_it]
print(f"mma: wait-full mean {mma_wait.mean():.0f}, sum/tile {mma_wait.sum()/tiles:.0f}; "
      f"tempty wait mean {(T[7,:tiles]-T[6,:tiles]).mean():.0f}")
ep_wait = T[1, 4000:4000+tiles] - T[0, 4000:4000+tiles]
ep_work = T[2, 4000:4000+tiles] - T[1, 4000:4000+tiles]
print(f"epilogue: wait-tfull mean {ep_wait.mean():.0f}, work mean {ep_work.mean():.0f}")
# a sample tile timeline (relative cycles)
t0 = T[0, per_tile * 5]
print("tile 5 stage timeline (rel cycles): issue/waitbeg/got/done")
for k in range(per_tile * 5, per_tile * 6):
    print(f"   it={k} b={k%nblk} load_issue={T[0,k]-t0 if T[0,k] else -1:7d} wait={T[1,k]-t0:7d} got={T[2,k]-t0:7d} done={T[3,k]-t0:7d} | mma wait={T[4,k]-t0:7d} full={T[5,k]-t0:7d}")
