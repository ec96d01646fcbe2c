from __future__ import annotations

import numpy as np
import torch


class RemoveSeenItems:
    """replay/models/nn/sequential/postprocessors/postprocessors.py:14-95: filters the user's WHOLE stored sequence.
    ``sequential`` is duck-typed (``get_sequence_by_query_id(query_ids, feature)``, ``schema``).  ``seen_tensor`` builds
    the padded [B, S] id matrix the fused kernel consumes (one host lookup per batch, as in the reference)."""

    def __init__(self, sequential):
        self._sequential = sequential
        self._candidates = None

    @property
    def candidates(self):
        return self._candidates

    @candidates.setter
    def candidates(self, c):
        self._candidates = c

    def seen_tensor(self, query_ids: torch.Tensor, device) -> torch.Tensor:
        name = self._sequential.schema.item_id_feature_name
        seqs = self._sequential.get_sequence_by_query_id(query_ids.flatten().cpu().numpy(), name)
        S = max(1, max(len(s) for s in seqs))
        out = np.full((len(seqs), S), -1, dtype=np.int64)
        for i, s in enumerate(seqs):
            out[i, : len(s)] = s
        return torch.from_numpy(out).to(device)

    def on_validation(self, query_ids, scores, ground_truth):
        """postprocessors.py:26-40"""
        return query_ids, self.on_prediction(query_ids, scores)[1], ground_truth

    def on_prediction(self, query_ids, scores):
        item_count = self._sequential.schema.item_id_features.item().cardinality
        seen = self.seen_tensor(query_ids, scores.device)
        ok = seen >= 0
        rows = torch.arange(scores.shape[0], device=scores.device).unsqueeze(1).expand_as(seen)
        out = scores.clone()
        if self._candidates is not None:
            full = torch.full((scores.shape[0], item_count), float("-inf"), device=scores.device, dtype=scores.dtype)
            full[:, self._candidates] = out
            full[rows[ok], seen[ok]] = float("-inf")
            return query_ids, full
        out[rows[ok], seen[ok]] = float("-inf")
        return query_ids, out
