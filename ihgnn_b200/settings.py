"""Mirror of the global settings the hot-path layers read
(/root/reference/Helpers/GlobalSettings.py:6-16 `Gsv`, :18-109 `Gs`).

When the reference is importable (drop-in use) its own `Gs` / `Gsv`
objects are used so that `Main.py`'s assignments are seen; otherwise this minimal mirror,
holding the reference defaults ("last assignment wins": transform = mean, dot-product
scoring, lambda 0.5, batch 100, 10 random negatives, lr 1e-3).
"""
try:  # pragma: no cover - exercised only next to the reference tree
    from Helpers.GlobalSettings import Gs, Gsv  # type: ignore
except Exception:  # reference not on sys.path: standalone mirror

    class Gsv:
        mean = "mean"
        activation = "activation"
        rnn = "rnn"
        graph_uqi = "uqi"
        graph_only_uq = "uq"
        graph_only_ui = "ui"
        graph_only_qi = "qi"

    class Gs:
        lambda_muq_for_hem = 0.5
        batch_size = 100
        learning_rate = 0.001
        embedding_size = 32
        weight_decay = 0
        random_negative_sample_size = 10
        non_random_negative_sample_size = 0
        graph_completeness = Gsv.graph_uqi

        class Query:
            transform = Gsv.mean

            @staticmethod
            def transform_activation():              # GlobalSettings.py:75-76 (last assignment wins: nn.ReLU)
                import torch.nn as nn
                return nn.ReLU()

        class Prediction:
            use_cosine_similarity = False

        class Debug:
            _calculate_embedding_info = False
            _calculate_highorder_info = False
