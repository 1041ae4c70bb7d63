"""Drop-in installation into the (unmodified) reference tree.

    import ihgnn_b200.install as inst
    inst.patch_reference("/path/to/IHGNN")     # before Main.py's imports run

or, as a launcher that leaves the reference untouched on disk:

    python -m ihgnn_b200.install /path/to/IHGNN/Main.py --device 0 --gnn IHGNN ...

`patch_reference` imports the reference's `Helpers.Graph`, `Models.*` and `Dataset` modules and
rebinds the hot-path classes -- PpsHyperGraph, Pps2DGraph, EmbeddingLayer, FeatureInteractor,
IHGNNLayer, HGCNLayer, GCNLayer, HemPredictionLayer -- to the CUDA-backed ones in every namespace that holds a
reference to them (`Models/__init__.py:12-24` name maps included), so `Main.py`, `RawGnn` and
`Srrl` run on top unchanged.  The reference imports `torch_sparse` and `dgl` unconditionally
(`Helpers/Torches.py:13-18`); they must be importable (the real packages, or the stubs under
oracle/stubs/ that the test-suite uses).
"""
from __future__ import annotations

import importlib
import runpy
import sys
from typing import Dict

_REPLACED = ("PpsHyperGraph", "Pps2DGraph", "EmbeddingLayer", "FeatureInteractor", "IHGNNLayer", "HGCNLayer",
             "GCNLayer", "HemPredictionLayer")


def replacement_classes() -> Dict[str, type]:
    from . import layers
    from .graph import Pps2DGraph, PpsHyperGraph
    out = {"PpsHyperGraph": PpsHyperGraph, "Pps2DGraph": Pps2DGraph}
    for name in _REPLACED:
        if name not in out:
            out[name] = getattr(layers, name)
    return out


def patch_reference(reference_dir: str = None, fast_eval: bool = True, fused_adam: bool = True) -> Dict[str, int]:
    """Rebind the hot-path classes inside the imported reference modules.  Returns, per class
    name, how many module attributes were rebound.  `fast_eval` also rebinds
    `Helpers.TrainTestHelper.test_and_get_avg_metrics` (:37-102) to the batched GPU ranking with the
    same signature and return value (`model.make_fast_test_and_get_avg_metrics`); Main.py picks it
    up through its `from Helpers.TrainTestHelper import ...`.  `fused_adam` makes `torch.optim.Adam(...)`
    (Main.py:192) return `ihgnn_b200.optim.FusedAdam` -- the same update, bit-identical with torch's fused
    CUDA Adam, as one kernel of this library -- whenever every parameter is a dense fp32 CUDA tensor and only
    lr / betas / eps / weight_decay are given; any other call goes to torch's own class."""
    if reference_dir and reference_dir not in sys.path:
        sys.path.insert(0, reference_dir)
    mods = [importlib.import_module(m) for m in (
        "Helpers.Graph", "Dataset", "Models.CommonLayers", "Models.EmbeddingLayers",
        "Models.GnnLayers", "Models.PredictionLayers", "Models.RawGnn", "Models.Srrl", "Models")]
    new = replacement_classes()
    old = {}
    for m in mods:
        for name in _REPLACED:
            cls = getattr(m, name, None)
            if isinstance(cls, type) and cls is not new[name]:
                old.setdefault(name, set()).add(cls)
    counts = {name: 0 for name in _REPLACED}
    for m in mods:
        for name in _REPLACED:
            if getattr(m, name, None) in old.get(name, ()):
                setattr(m, name, new[name])
                counts[name] += 1
        # dict/list registries built at import time (Models/__init__.py:15-24)
        for attr in ("parse_gnn_layer", "GnnLayerTypes"):
            reg = getattr(m, attr, None)
            if isinstance(reg, dict):
                for k, v in list(reg.items()):
                    for name in _REPLACED:
                        if v in old.get(name, ()):
                            reg[k] = new[name]
            elif isinstance(reg, list):
                for i, v in enumerate(reg):
                    for name in _REPLACED:
                        if v in old.get(name, ()):
                            reg[i] = new[name]
    if fast_eval:
        tth = importlib.import_module("Helpers.TrainTestHelper")
        if not getattr(tth.test_and_get_avg_metrics, "_ihgnn_b200", False):
            from .model import make_fast_test_and_get_avg_metrics
            metrics_cls = importlib.import_module("Helpers.Metrics").Metrics
            fast = make_fast_test_and_get_avg_metrics(tth.test_and_get_avg_metrics, metrics_cls)
            fast._ihgnn_b200 = True
            tth.test_and_get_avg_metrics = fast
    if fused_adam:
        _install_fused_adam()
    return counts


def _install_fused_adam() -> None:
    import torch
    original = torch.optim.Adam
    if getattr(original, "_ihgnn_b200", False):
        return
    from .optim import FusedAdam

    class Adam(original):                                  # isinstance(opt, torch.optim.Adam) keeps holding for torch's own
        _ihgnn_b200 = True

        def __new__(cls, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, **kw):
            params = list(params)
            flat = [p for g in params for p in g["params"]] if params and isinstance(params[0], dict) else params
            plain = not kw and not isinstance(lr, torch.Tensor) and flat and all(
                isinstance(p, torch.Tensor) and p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for p in flat)
            if plain:
                return FusedAdam(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
            return original(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, **kw)

    torch.optim.Adam = Adam


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m ihgnn_b200.install /path/to/IHGNN/Main.py [Main.py args...]")
    script = argv[0]
    import os
    patch_reference(os.path.dirname(os.path.abspath(script)))
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
