"""Loss selectors with the reference's names and constructor arguments (replay/nn/loss/{ce,bce}.py).  They carry no
computation: assigning one to ``SasRec.loss`` selects the fused CUDA head that implements it (full-catalog CE:
rp_ce_head_*; sampled heads: rp_sampled_head_*).  Single positive label per position (multi-positive: NotImplementedError,
as in the reference's CE)."""
from __future__ import annotations


class _LossSpec:
    kind = "ce"
    needs_negatives = False

    def engine_kwargs(self) -> dict:
        return {}

    # LossProto surface (replay/nn/loss/base.py:9-28): the fused path never calls a logits callback
    @property
    def logits_callback(self):
        return getattr(self, "_logits_callback", None)

    @logits_callback.setter
    def logits_callback(self, func):
        self._logits_callback = func


class CE(_LossSpec):
    """replay/nn/loss/ce.py:10-81: torch CrossEntropyLoss over the whole catalog, ``ignore_index`` = padding value."""

    def __init__(self, ignore_index: int = -100, **kwargs):
        if kwargs:
            raise NotImplementedError(f"CrossEntropyLoss options {sorted(kwargs)} are not supported by the fused head")
        self.ignore_index = ignore_index


class CESampled(_LossSpec):
    """replay/nn/loss/ce.py:146-249."""
    kind = "ce_sampled"
    needs_negatives = True

    def __init__(self, negative_labels_ignore_index: int = -100, **kwargs):
        if kwargs:
            raise NotImplementedError(f"CrossEntropyLoss options {sorted(kwargs)} are not supported by the fused head")
        self.negative_labels_ignore_index = negative_labels_ignore_index

    def engine_kwargs(self):
        return {"ignore_index": self.negative_labels_ignore_index}


class BCESampled(_LossSpec):
    """replay/nn/loss/bce.py:98-218."""
    kind = "bce_sampled"
    needs_negatives = True

    def __init__(self, log_epsilon: float = 1e-6, clamp_border: float = 100.0, negative_labels_ignore_index: int = -100):
        self.log_epsilon, self.clamp_border = log_epsilon, clamp_border
        self.negative_labels_ignore_index = negative_labels_ignore_index

    def engine_kwargs(self):
        return {"ignore_index": self.negative_labels_ignore_index, "log_eps": self.log_epsilon, "clamp": self.clamp_border}
