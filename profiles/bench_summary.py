"""Print a compact summary of a bench.py JSON line (file path argument)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["config"]["workload"], round(d["ms_per_step"], 3), "ms/step  value", f'{d["value"]:.4g}', d["unit"],
      " conv frac", round(d["roofline"]["conv_step"]["frac"], 3), " e2e ms", round(d["e2e"]["ms_per_step"], 3),
      " launches", d["gpu_launches"], " clocks", d["clocks"].get("sm_mhz"))
for k, v in d["kernels"].items():
    print("    ", k, v)
