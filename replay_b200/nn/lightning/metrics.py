"""Validation path: fused top-K feeding on-device ranking metrics.

Mirrors ``replay.nn.lightning.callback.ComputeMetricsCallback`` (metrics_callback.py:17-185) and the metric definitions of
``replay.metrics.torch_metrics_builder.TorchMetricsBuilder`` (torch_metrics_builder.py:268-393): recall / precision / ndcg /
map / mrr @k from the top-K item ids and a padded ground-truth matrix (padding < 0), accumulated on the device as sums and a
user count.  The top-K itself comes from the fused score + seen-filter + top-K kernel; metric arithmetic is a handful of torch
ops on [B, K] tensors (bookkeeping, not the hot path)."""
from __future__ import annotations

import torch

from ...compat import CallbackBase
from .postprocessor import SeenItemsFilter

_ALL = ("recall", "precision", "ndcg", "map", "mrr", "hitrate")


class RankingMetrics:
    def __init__(self, metrics=("recall", "ndcg", "map"), top_k=(10,)):
        for m in metrics:
            if m not in _ALL:
                raise ValueError(f"unsupported metric {m}; available: {_ALL}")
        self.metrics, self.top_k = tuple(metrics), tuple(sorted(top_k))
        self.max_k = max(self.top_k)
        self.names = [f"{m}@{k}" for k in self.top_k for m in self.metrics]
        self.reset()

    def reset(self):
        self._sum, self._n = None, 0

    def add_prediction(self, predictions: torch.Tensor, ground_truth: torch.Tensor):
        """predictions int64 [B, >=max_k]; ground_truth int64 [B, G] padded with negative values."""
        dev = predictions.device
        hits = (predictions[:, : self.max_k].unsqueeze(1) == ground_truth.unsqueeze(-1)).any(dim=1).float()
        gt = (ground_truth >= 0).sum(1).clamp(min=1)
        pos = torch.arange(2, 2 + self.max_k, device=dev).float()
        w_ndcg = 1.0 / torch.log2(pos)
        idcg = torch.cat([torch.zeros(1, device=dev), w_ndcg.cumsum(0)])
        w_map = 1.0 / torch.arange(1, 1 + self.max_k, device=dev).float()
        out = []
        for k in self.top_k:
            h, gk = hits[:, :k], gt.clamp(max=k)
            for m in self.metrics:
                if m == "recall":
                    v = h.sum(1) / gt
                elif m == "precision":
                    v = h.sum(1) / k
                elif m == "ndcg":
                    v = (h * w_ndcg[:k]).sum(1) / idcg[gk]
                elif m == "map":
                    v = (h * h.cumsum(1) * w_map[:k]).sum(1) / gk
                elif m == "hitrate":
                    v = (h.sum(1) > 0).float()
                else:  # mrr
                    ih = h * torch.arange(k, 0, -1, device=dev)
                    vals, idx = ih.max(dim=1)
                    v = (1.0 / (idx.masked_fill(vals == 0, -2) + 1).float()).clamp(min=0)
                out.append(v.sum())
        s = torch.stack(out)
        self._sum = s if self._sum is None else self._sum + s
        self._n += predictions.shape[0]

    def get_metrics(self) -> dict:
        assert self._n > 0
        return dict(zip(self.names, (self._sum / self._n).tolist()))


class ComputeMetricsCallback(CallbackBase):
    def __init__(self, metrics=("recall", "ndcg", "map"), ks=(10,), postprocessors=None, item_count=None,
                 ground_truth_column: str = "ground_truth"):
        self._builder = RankingMetrics(metrics, ks)
        self._postprocessors = postprocessors or []
        self._gt = ground_truth_column
        self.item_count = item_count

    def on_validation_epoch_start(self, trainer, pl_module):
        self._builder.reset()

    def on_validation_batch_end(self, trainer, pl_module, outputs, batch, batch_idx, dataloader_idx=0):
        model = getattr(pl_module, "model", None)
        k = self._builder.max_k
        from ...ops import MAX_FUSED_K

        if hasattr(model, "core") and k <= MAX_FUSED_K and all(isinstance(p, SeenItemsFilter) for p in self._postprocessors):
            seen = batch[self._postprocessors[0].seen_items_column] if self._postprocessors else None
            ids, _ = model.predict_topk(batch["feature_tensors"], batch["padding_mask"], k, seen, pl_module.candidates_to_score)
        else:
            logits = outputs["logits"]
            for p in self._postprocessors:
                logits = p.on_validation(batch, logits)
            ids = torch.topk(logits, k=k, dim=1).indices
        self._builder.add_prediction(ids, batch[self._gt])

    def on_validation_epoch_end(self, trainer, pl_module):
        m = self._builder.get_metrics()
        if hasattr(pl_module, "log_dict"):
            pl_module.log_dict(m, on_epoch=True, sync_dist=True)
        return m
