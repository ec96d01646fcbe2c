"""Validation / test path: fused top-K feeding on-device ranking metrics.

Mirrors ``replay.nn.lightning.callback.ComputeMetricsCallback`` (metrics_callback.py:17-240) and
``replay.metrics.torch_metrics_builder.TorchMetricsBuilder`` (torch_metrics_builder.py:93-393): recall / precision / ndcg / map /
mrr / novelty @k accumulated as sums + a user count, coverage@k from item histograms of the predictions and of the train sets,
plus ``hitrate`` (not in the reference builder; kept from round 1).  The top-K comes from the fused score + seen-filter + top-K
kernel whenever the model is engine-backed and the postprocessors are ``SeenItemsFilter``s (``[B, |I|]`` logits never exist);
the metric arithmetic is a handful of torch ops on ``[B, K]`` tensors - bookkeeping, not the hot path - and stays on the device
(one host transfer per epoch in ``get_metrics``; the reference calls ``.item()`` per metric and batch)."""
from __future__ import annotations

import torch

from ...compat import CallbackBase
from .postprocessor import SeenItemsFilter

_ALL = ("recall", "precision", "ndcg", "map", "mrr", "novelty", "coverage", "hitrate")
DEFAULT_METRICS = ("map", "ndcg", "recall")   # torch_metrics_builder.py:23-27
DEFAULT_KS = (1, 5, 10, 20)                   # torch_metrics_builder.py:29


class RankingMetrics:
    """``TorchMetricsBuilder(metrics, top_k, item_count)``: ``add_prediction(predictions, ground_truth, train=None)`` per batch,
    ``get_metrics()`` -> {"recall@10": ...}.  ``ground_truth`` / ``train`` are padded with values that are not item ids
    (negative), exactly as the reference requires."""

    def __init__(self, metrics=DEFAULT_METRICS, top_k=DEFAULT_KS, item_count: int | None = None):
        for m in metrics:
            if m not in _ALL:
                raise ValueError(f"unsupported metric {m}; available: {_ALL}")
        self.metrics = tuple(m for m in metrics if m != "coverage")
        self.need_coverage = "coverage" in metrics
        self.top_k = tuple(sorted(set(top_k)))
        self.max_k = max(self.top_k)
        self.item_count = item_count
        if self.need_coverage:
            assert item_count is not None, "For Coverage calculations item_count should be defined."
        # the reference's order: per k - recall, precision, ndcg, map, mrr, novelty (torch_metrics_builder.py:48-66)
        order = [m for m in ("recall", "precision", "ndcg", "map", "mrr", "novelty", "hitrate") if m in self.metrics]
        self._order = order
        self.names = [f"{m}@{k}" for k in self.top_k for m in order]
        self.reset()

    def reset(self):
        self._sum, self._n = None, 0
        self._train_hist = None
        self._pred_hist = {}

    def add_prediction(self, predictions: torch.Tensor, ground_truth: torch.Tensor, train: torch.Tensor | None = None):
        """predictions int64 [B, >= max_k]; ground_truth int64 [B, G]; train int64 [B, S] (novelty / coverage only)."""
        dev = predictions.device
        hits = (predictions[:, : self.max_k].unsqueeze(1) == ground_truth.unsqueeze(-1)).any(dim=1).float()
        train_hits = None
        if "novelty" in self._order:
            assert train is not None, "novelty needs the train items of every user"
            train_hits = (predictions[:, : self.max_k].unsqueeze(1) == train.unsqueeze(-1)).any(dim=1)
        gt = (ground_truth >= 0).sum(1).clamp(min=1)
        pos = torch.arange(2, 2 + self.max_k, device=dev).float()
        w_ndcg = 1.0 / torch.log2(pos)
        idcg = torch.cat([torch.zeros(1, device=dev), w_ndcg.cumsum(0)])
        w_map = 1.0 / torch.arange(1, 1 + self.max_k, device=dev).float()
        out = []
        for k in self.top_k:
            h, gk = hits[:, :k], gt.clamp(max=k)
            for m in self._order:
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
                elif m == "novelty":
                    v = (~train_hits[:, :k]).sum(1) / k
                else:  # mrr
                    ih = h * torch.arange(k, 0, -1, device=dev)
                    vals, idx = ih.max(dim=1)
                    v = (1.0 / (idx.masked_fill(vals == 0, -2) + 1).float()).clamp(min=0)
                out.append(v.sum())
        if out:
            s = torch.stack(out).double()
            self._sum = s if self._sum is None else self._sum + s
        self._n += predictions.shape[0]
        if self.need_coverage:  # _CoverageHelper (torch_metrics_builder.py:93-166): item histograms of predictions / train
            assert train is not None, "coverage needs the train items of every user"
            I = self.item_count
            if self._train_hist is None:
                self._train_hist = torch.zeros(I, device=dev)
                self._pred_hist = {k: torch.zeros(I, device=dev) for k in self.top_k}
            for k in self.top_k:
                p = predictions[:, :k].flatten()
                self._pred_hist[k] += torch.bincount(p[(p >= 0) & (p < I)], minlength=I).float()
            t = train.flatten()
            self._train_hist += torch.bincount(t[(t >= 0) & (t < I)], minlength=I).float()

    def get_metrics(self) -> dict:
        assert self._n > 0
        res = dict(zip(self.names, (self._sum / self._n).tolist())) if self._sum is not None else {}
        if self.need_coverage:
            seen = self._train_hist > 0
            n_train = int(seen.sum())
            for k in self.top_k:
                res[f"coverage@{k}"] = int((seen & (self._pred_hist[k] > 0)).sum()) / n_train
        return res


TorchMetricsBuilder = RankingMetrics  # the reference's name


class ComputeMetricsCallback(CallbackBase):
    """metrics_callback.py:17-240: validation AND test stages, one builder per dataloader, metrics history by epoch
    (``get_metrics(stage)``), ``state_dict`` / ``load_state_dict`` for checkpoints, ``train_column`` for novelty / coverage."""

    def __init__(self, metrics=None, ks=None, postprocessors=None, item_count=None, ground_truth_column: str = "ground_truth",
                 train_column: str = "train", verbose: bool = False):
        self._metrics = tuple(metrics or DEFAULT_METRICS)
        self._ks = tuple(ks or DEFAULT_KS)
        self._item_count = item_count
        self._postprocessors = postprocessors or []
        self._gt, self._train_column = ground_truth_column, train_column
        self._verbose = verbose
        self._builders: list[RankingMetrics] = [RankingMetrics(self._metrics, self._ks, item_count)]
        self._validation_metrics: dict[int, dict[str, float]] = {}
        self._test_metrics: dict[int, dict[str, float]] = {}
        self.item_count = item_count

    # ---- history / checkpoints (metrics_callback.py:72-100)
    def get_metrics(self, stage: str = "validate") -> dict:
        src = self._validation_metrics if stage == "validate" else self._test_metrics
        return {e: m.copy() for e, m in src.items()}

    def state_dict(self) -> dict:
        return {"validation_metrics": self._validation_metrics, "test_metrics": self._test_metrics}

    def load_state_dict(self, state_dict: dict) -> None:
        conv = lambda d: {int(e): {n: float(v) for n, v in m.items()} for e, m in d.items()}  # noqa: E731
        self._validation_metrics = conv(state_dict.get("validation_metrics", {}))
        self._test_metrics = conv(state_dict.get("test_metrics", {}))

    # ---- epoch / batch hooks
    def _epoch_start(self, n_loaders: int):
        self._builders = [RankingMetrics(self._metrics, self._ks, self._item_count) for _ in range(max(1, n_loaders))]

    @staticmethod
    def _n_loaders(trainer, attr):
        sizes = getattr(trainer, attr, None) if trainer is not None else None
        return len(sizes) if isinstance(sizes, (list, tuple)) else 1

    def on_validation_epoch_start(self, trainer, pl_module):
        self._epoch_start(self._n_loaders(trainer, "num_val_batches"))

    def on_test_epoch_start(self, trainer, pl_module):
        self._epoch_start(self._n_loaders(trainer, "num_test_batches"))

    def _batch_end(self, pl_module, outputs, batch, dataloader_idx):
        from ...ops import MAX_FUSED_K

        b = self._builders[dataloader_idx]
        model = getattr(pl_module, "model", None)
        k = b.max_k
        if hasattr(model, "core") and k <= MAX_FUSED_K and all(isinstance(p, SeenItemsFilter) for p in self._postprocessors):
            seen = batch[self._postprocessors[0].seen_items_column] if self._postprocessors else None
            ids, _ = model.predict_topk(batch["feature_tensors"], batch["padding_mask"], k, seen, pl_module.candidates_to_score)
        else:
            logits = outputs["logits"]
            for p in self._postprocessors:
                logits = p.on_validation(batch, logits)
            ids = torch.topk(logits, k=k, dim=1).indices
        b.add_prediction(ids, batch[self._gt], batch.get(self._train_column))

    def on_validation_batch_end(self, trainer, pl_module, outputs, batch, batch_idx, dataloader_idx=0):
        self._batch_end(pl_module, outputs, batch, dataloader_idx)

    def on_test_batch_end(self, trainer, pl_module, outputs, batch, batch_idx, dataloader_idx=0):
        self._batch_end(pl_module, outputs, batch, dataloader_idx)

    def _epoch_end(self, trainer, pl_module, is_validation: bool):
        if trainer is not None and getattr(trainer, "sanity_checking", False):
            return {}
        m = {}
        for i, b in enumerate(self._builders):
            if b._n == 0:
                continue
            suffix = "" if len(self._builders) == 1 else f"/dataloader_idx_{i}"
            m.update({k + suffix: v for k, v in b.get_metrics().items()})
        if hasattr(pl_module, "log_dict"):
            pl_module.log_dict(m, on_epoch=True, sync_dist=True)
        hist = self._validation_metrics if is_validation else self._test_metrics
        epoch = int(getattr(trainer, "current_epoch", len(hist)) or 0) if trainer is not None else len(hist)
        (self._validation_metrics if is_validation else self._test_metrics)[epoch] = dict(m)
        if self._verbose:
            print({k: round(v, 5) for k, v in m.items()})  # noqa: T201
        return m

    def on_validation_epoch_end(self, trainer, pl_module):
        return self._epoch_end(trainer, pl_module, True)

    def on_test_epoch_end(self, trainer, pl_module):
        return self._epoch_end(trainer, pl_module, False)
