"""ihgnn_b200 -- B200-native (sm_100a) implementation of IHGNN's interactive hypergraph
convolution hot path behind the reference's own layer API.

Importing the package does not load the CUDA library; the first kernel call does, and fails
loudly if libihgnn_b200.so is missing (there is no CPU or PyTorch fallback).
"""
from . import synth  # noqa: F401  (numpy-only)

__all__ = ["synth", "PpsHyperGraph", "Pps2DGraph", "GraphDataset", "EmbeddingLayer", "FeatureInteractor",
           "IHGNNLayer", "HGCNLayer", "GCNLayer", "HemPredictionLayer", "RawGnn", "FusedAdam"]


def __getattr__(name):
    if name in ("PpsHyperGraph", "Pps2DGraph"):
        from . import graph
        return getattr(graph, name)
    if name == "GraphDataset":
        from .dataset import GraphDataset
        return GraphDataset
    if name in ("EmbeddingLayer", "FeatureInteractor", "IHGNNLayer", "HGCNLayer", "GCNLayer", "HemPredictionLayer"):
        from . import layers
        return getattr(layers, name)
    if name == "RawGnn":
        from .model import RawGnn
        return RawGnn
    if name == "FusedAdam":
        from .optim import FusedAdam
        return FusedAdam
    raise AttributeError(name)
