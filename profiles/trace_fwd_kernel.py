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
n_it = int((T[5, :4000] > 0).sum())          # chunks seen by the MMA issuer
tiles = n_it // per_tile
span = T[5, :n_it].max() - T[4, 0]
print(f"block 0: {tiles} tiles, {n_it} chunks; span {span/1e3:.0f} kcycles = {span/max(tiles,1):.0f} cycles/tile, {span/max(n_it,1):.0f} cycles/chunk")
n_g = int((T[6, :2000] > 0).sum())           # chunks produced by group 0 (warp 0)
wait = T[2, :n_g] - T[1, :n_g]
prod = T[3, :n_g] - T[2, :n_g]
pub = T[6, :n_g] - T[3, :n_g]
period = np.diff(T[1, :n_g])
print(f"producer group 0 ({n_g} chunks): wait granule+A stage mean {wait.mean():.0f}; lds+products+split+st mean {prod.mean():.0f}; "
      f"wait::st+fence+arrive mean {pub.mean():.0f}; period median {np.median(period):.0f} (= 4 chunks)")
mw = T[5, :n_it] - T[4, :n_it]
iss = T[4, 1:n_it] - T[5, :n_it - 1]
print(f"mma: wait w_full + a_full mean {mw.mean():.0f}; issue 12 MMAs + commits median {np.median(iss):.0f}")
ep_wait = T[2, 6000:6000 + tiles] - T[1, 6000:6000 + tiles]
ep_work = T[3, 6000:6000 + tiles] - T[2, 6000:6000 + tiles]
print(f"epilogue: wait t_full mean {ep_wait.mean():.0f}, work mean {ep_work.mean():.0f}")

cta = np.zeros(320, dtype=np.int64)
lib.ihg_debug_read_cta_times.argtypes = [ctypes.c_void_p]
assert lib.ihg_debug_read_cta_times(cta.ctypes.data) == 0
cta = cta.reshape(160, 2)[:148]
t0 = cta[:, 0].min()
dur = cta[:, 1] - cta[:, 0]
print(f"per-CTA (ns): start spread {cta[:,0].max()-t0}, duration min {dur.min()} median {int(np.median(dur))} max {dur.max()}, "
      f"kernel span {cta[:,1].max()-t0}; block 0 duration {dur[0]}")
print("slowest CTAs:", np.argsort(dur)[-8:], dur[np.argsort(dur)[-8:]])
