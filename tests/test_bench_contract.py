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
               "--cpu-scale", "0.2")
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
    # config names the WORKLOAD (identical in both arms); the bounded sample is described in cpu_baseline
    assert d["config"] == {"workload": "amazon-small", "layers": 2, "dim": 64, "hyperedges": 100_000, "nodes": 35_000,
                           "interaction_order": 3, "scaling": "weak", "batch_rows": 1100}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "hyperedges" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_rank_workload():
    res = _run("--impl", "reference", "--workload", "rank", "--steps", "1", "--warmup", "1", "--cpu-rank-searches", "128")
    assert res.returncode == 0, res.stderr[-2000:]
    d = json.loads(res.stdout.strip())
    assert d["metric"] == "inference_ranking_searches_per_sec" and d["unit"] == "searches/s" and d["value"] > 0
    assert d["config"]["candidates"] == 1000 and d["config"]["searches"] == 131072 and d["cpu_baseline"]["kind"] == "port"


def test_default_workload_is_the_largest_single_gpu_config():
    res = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-scale", "0.004")
    assert res.returncode == 0, res.stderr[-2000:]
    d = json.loads(res.stdout.strip())
    assert d["config"]["workload"] == "cikm" and d["config"]["hyperedges"] == 5_000_000 and d["config"]["dim"] == 128


def test_reference_arm_other_ranks_exit_without_work():
    res = _run("--impl", "reference", "--workload", "amazon-small", "--steps", "1", "--warmup", "1",
               env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        return
    res = _run("--workload", "amazon-small", "--steps", "1", "--warmup", "1", "--no-cpu-baseline")
    assert res.returncode != 0 and "no CPU fallback" in res.stderr and res.stdout.strip() == ""
