"""Dev tool: in-kernel timeline of interact_bwd_slot_ts_kernel (block 0), instrumented build
(python -m ihgnn_b200.build --trace)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ihgnn_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "..", "..", "build", "libihgnn_trace.so")
from ihgnn_b200 import synth
from ihgnn_b200.dataset import GraphDataset
from ihgnn_b200.layers import FeatureInteractor
name = sys.argv[1] if len(sys.argv) > 1 else "amazon-full"
log = synth.make_workload(name); d = synth.WORKLOADS[name]["dim"]
ds = GraphDataset.from_search_log(log, "cuda:0")
fi = FeatureInteractor(ds, 3, d, d).to("cuda:0")
x = torch.randn(ds.node_count, d, device="cuda:0", requires_grad=True)
g = torch.randn(ds.graph.EdgeCount, d, device="cuda:0")
for _ in range(2):
    fi(x).backward(g)
torch.cuda.synchronize()
lib = _lib.lib()
lib.ihg_debug_read_trace_slot.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = np.zeros(1 << 16, dtype=np.int64)
assert lib.ihg_debug_read_trace_slot(buf.ctypes.data, buf.size) == 0
T = buf.reshape(8, 8192)
KC = d // 32
U = min(int((T[1, :2000] > 0).sum()), 1900 // KC)      # the probe rows hold 2000 entries per phase
tiles = U // KC
span = T[5, 2000:2000 + U].max() - T[1, 0]
print(f"{name} d={d}: block 0: {tiles} tiles, {U} units; span {span/1e3:.0f} kcycles = {span/max(tiles,1):.0f} cycles/tile, {span/max(U,1):.0f} cycles/unit")
def m(a): return f"{a.mean():.0f}"
print(f"gather: wait g_empty {m(T[0,2000:2000+U]-T[0,:U])}; ids + issue {m(T[0,4000:4000+U]-T[0,2000:2000+U])}; period {m(np.diff(T[0,:U]))}")
print(f"mma: wait t_empty {m(T[1,2000:2000+U]-T[1,:U])}; per chunk: wait w_full {m(T[3,:U*KC]-T[2,:U*KC])}, wait a_full {m(T[2,4000:4000+U*KC]-T[3,:U*KC])}; "
      f"chunk period {m(np.diff(T[2,:U*KC]))}; unit period {m(np.diff(T[1,:U]))}")
print(f"epilogue warp 4: wait g_full {m(T[4,2000:2000+U]-T[4,:U])}; wait t_full {m(T[4,4000:4000+U]-T[4,2000:2000+U])}; "
      f"ldtm+lds+math+sts {m(T[5,:U]-T[4,4000:4000+U])}; store {m(T[5,2000:2000+U]-T[5,:U])}; period {m(np.diff(T[4,:U]))}")
A = tiles * KC
print(f"A producer warp 0: cp.async wait {m(T[6,2000:2000+A]-T[6,:A])}; wait a_empty {m(T[6,4000:4000+A]-T[6,2000:2000+A])}; "
      f"lds+split+st+publish {m(T[7,:A]-T[6,4000:4000+A])}; period {m(np.diff(T[6,:A]))}")
t0 = T[1, 4 * KC]
print("timeline of tile 4 (rel cycles):")
for u in range(4 * KC, 5 * KC + 1):
    print(f"  unit {u}: G wait={T[0,u]-t0} got={T[0,2000+u]-t0} issued={T[0,4000+u]-t0} | M tempty_wait={T[1,u]-t0} ok={T[1,2000+u]-t0} | "
          f"E start={T[4,u]-t0} gfull={T[4,2000+u]-t0} tfull={T[4,4000+u]-t0} math_done={T[5,u]-t0} stored={T[5,2000+u]-t0}")
