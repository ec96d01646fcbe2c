from __future__ import annotations

import torch

from ...compat import CallbackBase
from .postprocessor import SeenItemsFilter


class TopItemsCallbackBase(CallbackBase):
    """replay/nn/lightning/callback/predictions_callback.py:29-163.  When the module's model is backed by the CUDA engine and
    the postprocessors are (only) ``SeenItemsFilter``s, scores, filter and top-K run as ONE fused kernel on the last hidden
    state - the ``[B, |I|]`` logits the reference materialises, clones, scatters into and re-reads are never built."""

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
        model = getattr(pl_module, "model", None)
        from ...ops import MAX_FUSED_K

        fusable = (hasattr(model, "core") and self._top_k <= MAX_FUSED_K
                   and all(isinstance(p, SeenItemsFilter) for p in self._postprocessors))
        if fusable:
            seen = batch[self._postprocessors[0].seen_items_column] if self._postprocessors else None
            ids, scores = model.predict_topk(batch["feature_tensors"], batch["padding_mask"], self._top_k, seen,
                                             pl_module.candidates_to_score)
        else:
            logits = outputs["logits"]
            for p in self._postprocessors:
                logits = p.on_prediction(batch, logits)
            scores, ids = torch.topk(logits, k=self._top_k, dim=1)
            if pl_module.candidates_to_score is not None:
                ids = torch.take(pl_module.candidates_to_score, ids)
        self._query_batches.append(batch[self.query_column])
        self._item_batches.append(ids)
        self._item_scores.append(scores)

    def get_result(self):
        return self._ids_to_result(torch.cat(self._query_batches), torch.cat(self._item_batches), torch.cat(self._item_scores))

    def _ids_to_result(self, query_ids, item_ids, item_scores):
        raise NotImplementedError


class TorchTopItemsCallback(TopItemsCallbackBase):
    def _ids_to_result(self, query_ids, item_ids, item_scores):
        return query_ids.flatten().cpu().long(), item_ids.cpu().long(), item_scores.cpu()


def _exploded_columns(query_ids, item_ids, item_scores):
    """One (query, item, score) row per recommendation - the frame the reference builds with ``explode`` - straight from the
    [B, K] tensors (no Python lists of per-row arrays)."""
    q = query_ids.flatten().cpu().numpy()
    k = item_ids.shape[1]
    return q.repeat(k), item_ids.cpu().numpy().reshape(-1), item_scores.cpu().numpy().reshape(-1)


class PandasTopItemsCallback(TopItemsCallbackBase):
    """predictions_callback.py:124-142"""

    def _ids_to_result(self, query_ids, item_ids, item_scores):
        import pandas as pd

        q, i, r = _exploded_columns(query_ids, item_ids, item_scores)
        return pd.DataFrame({self.query_column: q, self.item_column: i, self.rating_column: r})


class PolarsTopItemsCallback(TopItemsCallbackBase):
    """predictions_callback.py:145-163 (needs ``polars``; the import error of a missing package surfaces at ``get_result``)."""

    def _ids_to_result(self, query_ids, item_ids, item_scores):
        import polars as pl

        q, i, r = _exploded_columns(query_ids, item_ids, item_scores)
        return pl.DataFrame({self.query_column: q, self.item_column: i, self.rating_column: r})


class SparkTopItemsCallback(TopItemsCallbackBase):
    """predictions_callback.py:166-232: the result is a Spark DataFrame created by ``spark_session`` (integer query / item
    columns, double rating), one row per recommendation."""

    def __init__(self, top_k: int, query_column: str, item_column: str, rating_column: str, spark_session, postprocessors=None):
        super().__init__(top_k=top_k, query_column=query_column, item_column=item_column, rating_column=rating_column,
                         postprocessors=postprocessors)
        self.spark_session = spark_session

    def _ids_to_result(self, query_ids, item_ids, item_scores):
        from pyspark.sql.types import DoubleType, IntegerType, StructType

        schema = (StructType().add(self.query_column, IntegerType(), False).add(self.item_column, IntegerType(), False)
                  .add(self.rating_column, DoubleType(), False))
        q, i, r = _exploded_columns(query_ids, item_ids, item_scores)
        return self.spark_session.createDataFrame(data=list(zip(q.tolist(), i.tolist(), r.astype("float64").tolist())),
                                                  schema=schema)


class HiddenStatesCallback(CallbackBase):
    """predictions_callback.py:282-325: collects ``outputs["hidden_states"][hidden_state_index]`` of every predict batch."""

    def __init__(self, hidden_state_index: int):
        self._hidden_state_index = hidden_state_index
        self._embeddings_per_batch: list[torch.Tensor] = []

    def on_predict_epoch_start(self, trainer, pl_module):
        self._embeddings_per_batch.clear()

    def on_predict_batch_end(self, trainer, pl_module, outputs, batch, batch_idx, dataloader_idx=0):
        self._embeddings_per_batch.append(outputs["hidden_states"][self._hidden_state_index].detach().cpu())

    def get_result(self):
        return torch.cat(self._embeddings_per_batch)
