"""Dev tool: in-kernel timeline of node_linear_ts_kernel (block 0), instrumented build."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ihgnn_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "..", "..", "build", "libihgnn_trace.so")
from ihgnn_b200 import functional as F_
n, d = int(sys.argv[1]), int(sys.argv[2])
x = torch.randn(n, d, device="cuda:0"); w = torch.randn(1, d, d, device="cuda:0"); b = torch.randn(1, d, device="cuda:0")
for _ in range(3): y = F_.node_linear(x, w, bias=b)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); y = F_.node_linear(x, w, bias=b); e.record(); torch.cuda.synchronize()
print(f"node_linear {n} x {d}: {s.elapsed_time(e)*1e3:.0f} us (ideal {n*d*8/6.5e12*1e6:.0f} us)")
lib = _lib.lib()
lib.ihg_debug_read_trace_linear.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = np.zeros(8 * 4096, dtype=np.int64)
assert lib.ihg_debug_read_trace_linear(buf.ctypes.data, buf.size) == 0
T = buf.reshape(8, 4096)
KC = d // 32
tiles = int((T[6, 2048:3072] > 0).sum()); G = tiles * KC
t0 = T[7, 0]
print(f"block 0: {tiles} tiles; setup (barriers, TMEM alloc, W split) {T[7,1]-t0}; total {T[7,2]-t0} cycles = {(T[7,2]-t0)/max(tiles,1):.0f}/tile")
def m(a): return f"{a.mean():.0f}"
print(f"gather: wait g_empty {m(T[3,2048:2048+G]-T[3,:G])}; issue {m(T[4,:G]-T[3,2048:2048+G])}; period {m(np.diff(T[3,:G]))}")
P = int((T[2, :2048] > 0).sum())
print(f"producer group 0: wait g_full {m(T[1,:P]-T[0,:P])}; wait a_empty {m(T[1,2048:2048+P]-T[1,:P])}; lds+split+st+publish {m(T[2,:P]-T[1,2048:2048+P])}; period {m(np.diff(T[0,:P]))} (2 granules)")
print(f"mma: wait t_empty {m(T[5,2048:2048+tiles]-T[5,:tiles])}; tile period {m(np.diff(T[5,:tiles]))}")
print(f"epilogue: wait t_full {m(T[6,1024:1024+tiles]-T[6,:tiles])}; work {m(T[6,2048:2048+tiles]-T[6,1024:1024+tiles])}; period {m(np.diff(T[6,:tiles]))}")
print("first granules: gather issue done at", (T[4,:6]-t0).tolist(), " producer got g_full at", (T[1,:3]-t0).tolist(), " epilogue first t_full at", int(T[6,1024]-t0))
