"""Turn an `ncu --set full` report into the per-kernel summary + DRAM-traffic json kept under profiles/.

    python profiles/ncu_extract.py gpurun_out/prof.ncu-rep profiles/r02_ncu_full_cikm.txt \
        [--traffic profiles/r02_ncu_traffic_cikm.json --workload cikm --steps-captured 1] [--note "..."]

(`--steps-captured`: the report was taken with `ncu --profile-from-start off ... bench.py --profile-step`, i.e. it
holds the launches of exactly one eager conv step; bench.py reads `dram_bytes_per_step` for roofline.dram_step.)

Reads the report with `ncu -i <rep> --page raw --csv` (works without a GPU) and averages every
metric of interest over the captured launches of each kernel.
"""
import argparse
import collections
import csv
import io
import json
import re
import subprocess

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]
_TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
_TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    return name.replace("ihg::", "").replace("(anonymous namespace)::", "")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--traffic")
    ap.add_argument("--workload", default="amazon-full")
    ap.add_argument("--note", default="")
    ap.add_argument("--how", default="--set full", help="how the report was collected (goes into the header line)")
    ap.add_argument("--steps-captured", type=int, default=0,
                    help="the report holds exactly this many whole steps: also write dram_bytes_per_step")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(header)}
    per = collections.OrderedDict()
    for r in data:
        k = short(r[col["Kernel Name"]])
        d = per.setdefault(k, collections.defaultdict(list))
        for m in METRICS:
            if m in col and r[col[m]] not in ("", "n/a"):
                d[m].append(float(r[col[m]].replace(",", "")))
    lines = [f"ncu {a.how} --clock-control none; per-kernel means over the captured launches. {a.note}".rstrip(), ""]
    traffic = {}
    total_bytes, total_launches = 0.0, 0
    for k, d in per.items():
        n = len(d["gpu__time_duration.sum"])
        lines.append(f"{k}  ({n} launches)")
        for m in METRICS:
            if d[m]:
                lines.append(f"    {m:<88} {sum(d[m]) / len(d[m]):14.3f} {units[col[m]]}")
        lines.append("")
        if d["dram__bytes_read.sum"]:
            ur, uw = units[col["dram__bytes_read.sum"]], units[col["dram__bytes_write.sum"]]
            traffic[k] = (sum(d["dram__bytes_read.sum"]) / n) * _TO_BYTES.get(ur, 1.0) + \
                         (sum(d["dram__bytes_write.sum"]) / n) * _TO_BYTES.get(uw, 1.0)
            total_bytes += traffic[k] * n
            total_launches += n
    with open(a.out, "w") as f:
        f.write("\n".join(lines))
    if a.traffic:
        with open(a.traffic, "w") as f:
            doc = {"workload": a.workload, "source": a.out, "dram_bytes_per_launch": traffic}
            if a.steps_captured > 0:
                doc["dram_bytes_per_step"] = total_bytes / a.steps_captured
                doc["launches_per_step"] = total_launches / a.steps_captured
            json.dump(doc, f, indent=1)
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
