"""Pre-/in-field helper layers used by the reference's configs (torchdyn/nn/node_layers.py): plain torch modules that
sit either side of the integration path (callers of the hot path)."""
from .node_layers import Augmenter, DataControl, DepthCat

__all__ = ["Augmenter", "DepthCat", "DataControl"]
