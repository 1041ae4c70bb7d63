"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the
required keys, and the GPU arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

from conftest import REPO


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), *args], capture_output=True, text=True,
                          cwd=REPO, env=e, timeout=600)


def test_reference_arm_prints_one_contract_line():
    res = _run("--impl", "reference", "--workload", "amazon-small", "--steps", "1", "--warmup", "1",
               "--cpu-sample-edges", "20000")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "hypergraph_conv_hyperedge_layers_per_sec_fwd_bwd"
    for k in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"] == "amazon-small" and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "hyperedges" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_work():
    res = _run("--impl", "reference", "--workload", "amazon-small", "--steps", "1", "--warmup", "1",
               env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        return
    res = _run("--workload", "amazon-small", "--steps", "1", "--warmup", "1", "--no-cpu-baseline")
    assert res.returncode != 0 and "no CPU fallback" in res.stderr and res.stdout.strip() == ""
