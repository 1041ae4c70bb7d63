"""Host-side readers of the reference's on-disk dataset format (CPU): round trip through
synth.write_reference_files, and -- where /root/reference is present -- equality with the reference's own
GraphDataset / TestSearchLogDataLoader parse of the same files."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

from conftest import REPO
from ihgnn_b200 import synth
from ihgnn_b200.dataset import read_reference_files, read_search_logs, read_test_logs

REFERENCE = "/root/reference"


@pytest.mark.parametrize("shape", ["amazon", "cikm"])
def test_read_reference_files_round_trip(tmp_path, shape):
    log = synth.make_search_log(60, 25, 90, 400, 40, shape=shape, seed=3, with_negatives=True)
    synth.write_reference_files(log, str(tmp_path))
    d = read_reference_files(str(tmp_path))
    assert (d["user_count"], d["query_count"], d["item_count"], d["vocab_size"]) == (60, 25, 90, 40)
    assert np.array_equal(d["pos_user"], log.pos_user) and np.array_equal(d["pos_query"], log.pos_query)
    assert np.array_equal(d["pos_item"], log.pos_item)
    words, offsets = log.bag_inputs()
    assert np.array_equal(d["bag_words"], words) and np.array_equal(d["bag_offsets"], offsets)
    assert np.array_equal(d["log_ptr"], log.log_ptr) and np.array_equal(d["log_items"], log.log_items)
    assert np.array_equal(d["log_flags"], log.log_flags)
    tests = read_test_logs(os.path.join(str(tmp_path), "test_data.csv"))
    assert len(tests) == 8 and all(len(t[2]) == 1 and t[3] is None and t[4] is True for t in tests)
    u, q, ptr, items, flags = read_search_logs(os.path.join(str(tmp_path), "valid_data.csv"))
    assert u.shape == q.shape == (8,) and ptr[-1] == items.shape[0] == flags.shape[0]


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree not present")
def test_readers_match_the_reference_parse(tmp_path):
    """The unmodified reference parses the same files into the same positives, EmbeddingBag inputs and
    test logs (Dataset.py:141-212, :297-318)."""
    import subprocess
    log = synth.make_search_log(40, 18, 60, 320, 30, shape="cikm", seed=16, with_negatives=True)
    synth.write_reference_files(log, str(tmp_path))
    code = r'''
import sys, json, contextlib, io
sys.dont_write_bytecode = True
sys.path.insert(0, "%(repo)s/oracle/stubs"); sys.path.insert(0, "%(ref)s")
import torch
with contextlib.redirect_stdout(io.StringIO()):
    from Helpers.GlobalSettings import Gs, Gsv
    Gs.graph_completeness = Gsv.graph_uqi
    from Helpers.IOHelper import IOHelper
    IOHelper.warned_about_cannot_log = True
    from Helpers.Graph import PpsHyperGraph
    from Dataset import GraphDataset, TestSearchLogDataLoader
    ds = GraphDataset("%(d)s/graph_info.txt", "%(d)s/queries_multihot.txt", "%(d)s/train_data.csv", PpsHyperGraph, 10, 0, torch.device("cpu"))
    tl = TestSearchLogDataLoader("%(d)s/test_data.csv", ds, torch.device("cpu"))
print(json.dumps({"pos": [list(p.uqif()[:3]) for p in ds.pos_interactions],
                  "words": ds.queries_for_embeddingbag.tolist(), "offsets": ds.queries_offset_for_embeddingbag.tolist(),
                  "tests": [[l[0], l[1], list(l[2])] for l in tl.logs]}))
''' % {"repo": REPO, "ref": REFERENCE, "d": str(tmp_path)}
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path))
    assert res.returncode == 0, res.stderr[-2000:]
    import json
    ref = json.loads(res.stdout.strip().splitlines()[-1])
    d = read_reference_files(str(tmp_path))
    pos = np.asarray(ref["pos"], dtype=np.int64)
    assert np.array_equal(d["pos_user"], pos[:, 0]) and np.array_equal(d["pos_query"], pos[:, 1])
    assert np.array_equal(d["pos_item"], pos[:, 2])
    assert d["bag_words"].tolist() == ref["words"] and d["bag_offsets"].tolist() == ref["offsets"]
    mine = read_test_logs(os.path.join(str(tmp_path), "test_data.csv"))
    assert [[t[0], t[1], t[2]] for t in mine] == ref["tests"]


def test_metrics_at_10_matches_reference_metrics(golden):
    """Host metric arithmetic of the batched evaluation against the values the reference's own
    Metrics.calculate_on_all_items produced (tests/golden, oracle/gen_golden.py)."""
    from ihgnn_b200.model import metrics_at_10
    for b in range(golden["ref64.rank_top10"].shape[0]):
        inter = [int(x) for x in golden["rank.interacted"][b] if x >= 0]
        got = metrics_at_10(golden["ref64.rank_top10"][b], inter)
        assert np.allclose(got, golden["ref64.rank_metrics"][b], rtol=0, atol=1e-12), (b, got)


def test_logged_negative_lists_match_reference_dict(tmp_path):
    """`neg_items_for_user_query_pair` (Dataset.py:196-209): brute-force dict, and -- where the reference
    tree is present -- the dict the reference's own GraphDataset builds from the same files."""
    import json
    import subprocess
    from ihgnn_b200.dataset import logged_negative_lists
    log = synth.make_search_log(40, 18, 60, 320, 30, shape="cikm", seed=16, with_negatives=True)
    pp, nptr, nit = logged_negative_lists(log.log_user, log.log_query, log.log_ptr, log.log_items, log.log_flags,
                                          log.pos_user, log.pos_query, 18)
    want = {}
    for k in range(len(log.log_user)):
        a, b = int(log.log_ptr[k]), int(log.log_ptr[k + 1])
        lst = want.setdefault((int(log.log_user[k]), int(log.log_query[k])), [])
        lst.extend(int(it) for it, fl in zip(log.log_items[a:b], log.log_flags[a:b]) if fl <= 0)
    assert nptr[-1] == nit.shape[0] == sum(len(v) for v in want.values())
    for e in range(len(log.pos_user)):
        got = nit[nptr[pp[e]]:nptr[pp[e] + 1]].tolist()
        assert got == want[(int(log.pos_user[e]), int(log.pos_query[e]))]
    if not os.path.isdir(REFERENCE):
        return
    synth.write_reference_files(log, str(tmp_path))
    code = r'''
import sys, json, contextlib, io
sys.dont_write_bytecode = True
sys.path.insert(0, "%(repo)s/oracle/stubs"); sys.path.insert(0, "%(ref)s")
import torch
with contextlib.redirect_stdout(io.StringIO()):
    from Helpers.GlobalSettings import Gs, Gsv
    Gs.graph_completeness = Gsv.graph_uqi
    from Helpers.IOHelper import IOHelper
    IOHelper.warned_about_cannot_log = True
    from Helpers.Graph import PpsHyperGraph
    from Dataset import GraphDataset
    ds = GraphDataset("%(d)s/graph_info.txt", "%(d)s/queries_multihot.txt", "%(d)s/train_data.csv", PpsHyperGraph, 8, 2, torch.device("cpu"))
print(json.dumps([[k[0], k[1], v] for k, v in ds.neg_items_for_user_query_pair.items()]))
''' % {"repo": REPO, "ref": REFERENCE, "d": str(tmp_path)}
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path))
    assert res.returncode == 0, res.stderr[-2000:]
    ref = {(u, q): v for u, q, v in json.loads(res.stdout.strip().splitlines()[-1])}
    assert ref == want
