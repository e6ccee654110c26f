"""istvt_b200 — B200-native (sm_100a) implementation of the ISTVT forward hot path.

The directory name (`2023-tifs-istvt_b200`) is not a Python identifier; import it with

    import importlib
    istvt = importlib.import_module("2023-tifs-istvt_b200")

or put this directory itself on `sys.path` to shadow the reference's `network` package
(`from network.models import model_selection` then resolves to the B200 implementation).
"""
from . import _lib, ops  # noqa: F401
from .engine import ISTVTEngine, entry_flow_features  # noqa: F401
from .stream import ClipStream  # noqa: F401
from .train import Trainer  # noqa: F401
from .relevance import LRP, relevance_maps  # noqa: F401
from .graph import GraphedForward  # noqa: F401
from .network.models import TransferModel, model_selection  # noqa: F401
from .network.vivit.vivit import DSTTr, STTransformer, Transformer, VanillaTr, ViViT, XceptionVidTr  # noqa: F401
from .network.vivit.module import Attention, TemporalOnlyAttention  # noqa: F401

ISTVT = XceptionVidTr

__all__ = ["ISTVT", "XceptionVidTr", "DSTTr", "STTransformer", "TransferModel", "model_selection", "ISTVTEngine",
           "entry_flow_features", "ops", "Transformer", "ViViT", "VanillaTr", "Attention", "TemporalOnlyAttention", "ClipStream", "Trainer", "LRP", "relevance_maps", "GraphedForward"]
