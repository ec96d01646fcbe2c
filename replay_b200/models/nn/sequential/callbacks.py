from __future__ import annotations

import torch

from ....compat import CallbackBase
from .postprocessors import RemoveSeenItems


class BasePredictionCallback(CallbackBase):
    """replay/models/nn/sequential/callbacks/prediction_callbacks.py:30-120.  With a ``replay_b200`` module and only
    ``RemoveSeenItems`` postprocessors the scores / filter / top-K run fused on the device."""

    def __init__(self, top_k: int, query_column: str, item_column: str, rating_column: str = "rating", postprocessors=None):
        self.query_column, self.item_column, self.rating_column = query_column, item_column, rating_column
        self._top_k = top_k
        self._postprocessors = postprocessors or []
        self._query_batches, self._item_batches, self._item_scores = [], [], []

    def on_predict_epoch_start(self, trainer, pl_module):
        self._query_batches.clear(); self._item_batches.clear(); self._item_scores.clear()
        for p in self._postprocessors:
            p.candidates = pl_module.candidates_to_score

    def on_predict_batch_end(self, trainer, pl_module, outputs, batch, batch_idx, dataloader_idx=0):
        query_ids = batch["query_id"]
        fusable = hasattr(pl_module, "predict_topk") and all(isinstance(p, RemoveSeenItems) for p in self._postprocessors)
        if fusable:
            dev = batch["padding_mask"].device
            seen = self._postprocessors[0].seen_tensor(query_ids, dev) if self._postprocessors else None
            ids, scores = pl_module.predict_topk(batch, self._top_k, seen, pl_module.candidates_to_score)
        else:
            scores_full = outputs
            for p in self._postprocessors:
                query_ids, scores_full = p.on_prediction(query_ids, scores_full)
            scores, ids = torch.topk(scores_full, k=self._top_k, dim=1)
        self._query_batches.append(query_ids)
        self._item_batches.append(ids)
        self._item_scores.append(scores)

    def get_result(self):
        return self._ids_to_result(torch.cat(self._query_batches), torch.cat(self._item_batches), torch.cat(self._item_scores))

    def _ids_to_result(self, q, i, s):
        raise NotImplementedError


class TorchPredictionCallback(BasePredictionCallback):
    def _ids_to_result(self, q, i, s):
        return q.flatten().cpu().long(), i.cpu().long(), s.cpu()


class PandasPredictionCallback(BasePredictionCallback):
    def _ids_to_result(self, q, i, s):
        import pandas as pd

        k = i.shape[1]
        return pd.DataFrame({self.query_column: q.flatten().cpu().numpy().repeat(k),
                             self.item_column: i.cpu().numpy().reshape(-1), self.rating_column: s.cpu().numpy().reshape(-1)})
