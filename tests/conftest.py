"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path.

`-m "not gpu"` runs here on CPU (oracle vs golden vectors, host logic, C-ABI symbol
checks, gloo world_size-2 tests); `-m gpu` are the parity tests proper, run on a B200.
"""
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


@pytest.fixture(params=golden_names())
def golden(request, monkeypatch):
    import numpy as np
    with np.load(os.path.join(GOLDEN_DIR, request.param + ".npz")) as z:
        data = {k: z[k] for k in z.files}
    data["__name__"] = request.param
    # fixtures generated with Gs.Prediction.use_cosine_similarity = True run the product with it too
    from ihgnn_b200 import settings
    monkeypatch.setattr(settings.Gs.Prediction, "use_cosine_similarity", bool(data.get("cfg.cosine", False)))
    act = str(data["cfg.query_activation"]) if "cfg.query_activation" in data else ""
    monkeypatch.setattr(settings.Gs.Query, "transform", settings.Gsv.activation if act else settings.Gsv.mean)
    return data
