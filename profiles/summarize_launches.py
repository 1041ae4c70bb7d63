"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        text = f.read()
    agg = collections.OrderedDict()
    if text.startswith("id,kernel,duration("):
        # compact list written by profiles/scripts/r02_run_ncu.sh: id,kernel,duration(<unit>)
        unit = text[text.index("(") + 1:text.index(")")]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit.replace("second", "s"), 1.0)
        for line in text.splitlines()[1:]:
            _id, rest = line.split(",", 1)
            name, dur = rest.rsplit(",", 1)
            a = agg.setdefault(name.replace("ihg::", "").replace("<unnamed>::", ""), [0, 0.0])
            a[0] += 1
            a[1] += float(dur) * scale
        lines = []
    else:
        lines = [l for l in text.splitlines(True) if l.startswith('"')]
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit.startswith("n") else (v * 1e3 if unit.startswith("m") else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'total us':>12} {'calls':>6} {'share':>7}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:12.1f} {n:6d} {100 * t / tot:6.1f}%  {k[:110]}")
    print(f"{tot:12.1f} us total over {sum(a[0] for a in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1])
