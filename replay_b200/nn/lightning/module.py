from __future__ import annotations

import inspect

import torch

from ...compat import LightningModuleBase


class OptimizerFactory:
    """replay/nn/lightning/optimizer.py:24-60 - Adam(lr 1e-3, betas (0.9, 0.98)) by default."""

    def __init__(self, optimizer: str = "adam", learning_rate: float = 0.001, weight_decay: float = 0.0,
                 betas: tuple = (0.9, 0.98)):
        if optimizer != "adam" or weight_decay != 0.0:
            raise NotImplementedError("the fused B200 path implements Adam without weight decay (the reference default)")
        self.learning_rate, self.betas = learning_rate, betas

    def create(self, parameters):
        return torch.optim.Adam(parameters, lr=self.learning_rate, betas=self.betas)


class LightningModule(LightningModuleBase):
    """replay/nn/lightning/module.py:13-123.  ``fused_optimizer=True`` (default) runs forward+backward+Adam inside the CUDA
    engine (manual optimisation); with False the loss goes through autograd and the optimizer from ``optimizer_factory``."""

    def __init__(self, model, optimizer_factory: OptimizerFactory | None = None, lr_scheduler_factory=None,
                 fused_optimizer: bool = True):
        super().__init__()
        self.save_hyperparameters(ignore=["model"])
        self.model = model
        self._optimizer_factory = optimizer_factory or OptimizerFactory()
        self._lr_scheduler_factory = lr_scheduler_factory
        self._candidates_to_score = None
        self.fused_optimizer = fused_optimizer
        if fused_optimizer:
            self.automatic_optimization = False
        self._sig = set(inspect.signature(model.forward).parameters)

    def forward(self, batch: dict):
        if "candidates_to_score" in self._sig and self._candidates_to_score is not None and not self.model.training:
            batch = {**batch, "candidates_to_score": self._candidates_to_score}
        return self.model(**{k: v for k, v in batch.items() if k in self._sig})

    def training_step(self, batch: dict, batch_idx: int = 0):
        if self.fused_optimizer and hasattr(self.model, "core"):
            core = self.model.core
            lab, tm = batch["positive_labels"], batch["target_padding_mask"]
            lab = lab[..., 0] if lab.dim() == 3 else lab
            tm = tm[..., 0] if tm.dim() == 3 else tm
            spec = getattr(self.model, "loss", None)
            neg = batch.get("negative_labels") if getattr(spec, "needs_negatives", False) else None
            if getattr(spec, "needs_negatives", False) and neg is None:
                raise ValueError(f"{type(spec).__name__} needs `negative_labels` in the batch")
            loss = core.fused_step(batch["feature_tensors"][core.item_feature], batch["padding_mask"], lab, tm,
                                   lr=self._optimizer_factory.learning_rate, negatives=neg)
        else:
            loss = self(batch)["loss"]
        self.log("train_loss", loss, on_step=True, on_epoch=True, prog_bar=True, sync_dist=True)
        return loss

    def predict_step(self, batch: dict, batch_idx: int = 0, dataloader_idx: int = 0):
        self.model.eval()
        return self(batch)

    def configure_optimizers(self):
        opt = self._optimizer_factory.create(self.model.parameters())
        if self._lr_scheduler_factory is None:
            return opt
        return [opt], [self._lr_scheduler_factory.create(opt)]

    @property
    def candidates_to_score(self):
        return self._candidates_to_score

    @candidates_to_score.setter
    def candidates_to_score(self, candidates):
        if candidates is not None:
            if not (isinstance(candidates, torch.Tensor) and candidates.dtype == torch.long and candidates.dim() == 1):
                raise ValueError("candidates_to_score must be a 1-D torch.LongTensor")
            if candidates.unique().numel() != candidates.numel():
                raise ValueError("candidates_to_score must contain unique item ids")  # module.py:118-123
        self._candidates_to_score = candidates
