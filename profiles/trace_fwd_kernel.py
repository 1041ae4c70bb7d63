"""Dev tool: in-kernel timeline of feature_interact_fwd_ts_kernel (block 0) using an instrumented
build of the library (build/libihgnn_trace.so, python -m ihgnn_b200.build --trace: -DIHG_TRACE clock64
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
G = tiles * KC
span = T[3, :n_it].max() - T[0, 0]
print(f"block 0: {tiles} tiles, {n_it} chunks; span {span/1e3:.0f} kcycles = {span/max(tiles,1):.0f} cycles/tile, {span/max(n_it,1):.0f} cycles/chunk")
cpw = T[6, :G] - T[0, :G]
barw = T[7, :G] - T[6, :G]
print(f"producer granule: cp.async wait mean {cpw.mean():.0f}, barrier mean {barw.mean():.0f}")
iss = T[4, 7000:7000 + G] - T[7, :G]
ldsw = T[5, 7000:7000 + G] - T[4, 7000:7000 + G]
print(f"producer granule: issue next gathers + id prefetch mean {iss.mean():.0f}; lds u,q,i mean {ldsw.mean():.0f}")
n6 = min(n_it, 3000)
spl = T[6, 1000:1000 + n6] - np.where(np.arange(n6) % nblk == 0, T[5, 7000 + np.arange(n6) // nblk], T[3, np.maximum(np.arange(n6) - 1, 0)])
pub = T[1, :n6] - T[6, 1000:1000 + n6]
print(f"producer chunk: products+split mean {spl.mean():.0f} (median {np.median(spl):.0f}); publish previous (wait::st, fence, arrive) mean {pub.mean():.0f}")
gi = np.arange(G)
first_it = gi * nblk
lds = T[1, first_it] - T[7, :G]
print(f"producer: barrier -> first chunk ready (issue next gathers + lds + split) mean {lds.mean():.0f}")
wait_empty = T[2, :n_it] - T[1, :n_it]
st = T[3, :n_it] - T[2, :n_it]
gap = T[1, 1:n_it] - T[3, :n_it - 1]
print(f"producer chunk: wait a_empty mean {wait_empty.mean():.0f}; st+wait::st+arrive mean {st.mean():.0f}; compute gap to next wait mean {np.median(gap):.0f} (median)")
ww = T[0, 4096:4096 + n_it] - T[4, :n_it]
aw = T[5, :n_it] - T[0, 4096:4096 + n_it]
issue = T[4, 1:n_it] - T[5, :n_it - 1]
print(f"mma: wait w_full mean {ww.mean():.0f}; wait a_full mean {aw.mean():.0f}; issue 12 MMAs + commits (median) {np.median(issue):.0f}")
ep_wait = T[2, 6000:6000 + tiles] - T[1, 6000:6000 + tiles]
ep_work = T[3, 6000:6000 + tiles] - T[2, 6000:6000 + tiles]
print(f"epilogue: wait t_full mean {ep_wait.mean():.0f}, work mean {ep_work.mean():.0f}")
t0 = T[0, KC * 5]
print("tile 5 timeline (rel cycles)")
for g in range(KC * 5, KC * 6):
    print(f" granule {g}: top={T[0,g]-t0} cpwait_done={T[6,g]-t0} barrier_done={T[7,g]-t0}")
    for k in range(g * nblk, (g + 1) * nblk):
        print(f"   it={k} b={k%nblk} P: wait={T[1,k]-t0:7d} got={T[2,k]-t0:7d} arrived={T[3,k]-t0:7d} | M: wait={T[4,k]-t0:7d} w_ok={T[0,4096+k]-t0:7d} a_ok={T[5,k]-t0:7d}")
