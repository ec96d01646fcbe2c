from __future__ import annotations

import torch


class SeenItemsFilter:
    """replay/nn/lightning/postprocessor/seen_items.py:8-83: scores of already seen items become -inf.  Ids outside
    [0, item_count) in ``batch[seen_items_column]`` are padding.  The top-items callbacks recognise this class and fuse the
    filter into the score + top-K kernel instead of calling ``on_prediction`` on materialised logits."""

    def __init__(self, item_count: int, seen_items_column: str = "seen_ids"):
        self.item_count = item_count
        self.seen_items_column = seen_items_column
        self._candidates = None

    @property
    def candidates(self):
        return self._candidates

    @candidates.setter
    def candidates(self, c):
        self._candidates = c

    def _compute(self, batch, logits):
        seen = batch[self.seen_items_column]
        ok = (seen >= 0) & (seen < self.item_count)
        rows = torch.arange(logits.shape[0], device=logits.device).unsqueeze(1).expand_as(seen)
        out = logits.detach().clone()
        if self._candidates is None:
            out[rows[ok], seen[ok]] = float("-inf")
            return out
        full = torch.full((logits.shape[0], self.item_count), float("-inf"), device=logits.device, dtype=logits.dtype)
        full[:, self._candidates] = out
        full[rows[ok], seen[ok]] = float("-inf")
        return full[:, self._candidates]

    def on_validation(self, batch, logits):
        return self._compute(batch, logits)

    def on_prediction(self, batch, logits):
        return self._compute(batch, logits)
