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


class LazyInferenceOutput(dict):
    """``InferenceOutput`` (replay/nn/output.py) whose ``logits`` [B, |I|] and ``hidden_states`` are computed on first access.
    ``predict_step`` / ``validation_step`` return it, so a callback that runs the fused score + seen-filter + top-K head
    (``TopItemsCallbackBase``, ``ComputeMetricsCallback``) never makes the body run twice nor the [B, |I|] fp32 scores get
    written (8 GB per 4096 users at 500 K items); a callback that does read ``outputs["logits"]`` gets the reference's tensor."""

    def __init__(self, compute):
        super().__init__()
        self._compute = compute

    def _fill(self):
        if self._compute is not None:
            self._inner, self._compute = self._compute(), None
            super().update(self._inner)

    def __getitem__(self, k):
        self._fill()
        if not super().__contains__(k):  # the model's own output may compute entries lazily as well (hidden_states)
            super().__setitem__(k, self._inner[k])
        return super().__getitem__(k)

    def __contains__(self, k):
        return k in ("logits", "hidden_states") or super().__contains__(k)

    def get(self, k, default=None):
        self._fill()
        return super().get(k, default)

    def keys(self):
        self._fill()
        return super().keys()

    def items(self):
        self._fill()
        return super().items()

    def values(self):
        self._fill()
        return super().values()

    @property
    def materialised(self) -> bool:
        return self._compute is None


class LightningModule(LightningModuleBase):
    """replay/nn/lightning/module.py:13-123.  ``fused_optimizer=True`` (default) runs forward+backward+Adam inside the CUDA
    engine (manual optimisation; under ``torch.distributed`` the flat gradient is all-reduced before Adam, which is what
    Lightning's DDP does for the reference); with False the loss goes through autograd and the optimizer from
    ``optimizer_factory``.  In fused mode the learning rate of every step is read from the optimizer Lightning configured
    (so an lr scheduler takes effect), ``betas`` come from the factory."""

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
        core = getattr(model, "core", None)
        if core is not None:
            core.adam_betas = tuple(getattr(self._optimizer_factory, "betas", (0.9, 0.98)))

    def forward(self, batch: dict):
        if "candidates_to_score" in self._sig and self._candidates_to_score is not None and not self.model.training:
            batch = {**batch, "candidates_to_score": self._candidates_to_score}
        return self.model(**{k: v for k, v in batch.items() if k in self._sig})

    # ---- checkpoints: the reference's keys are ``model.`` + the model's own; torch's recursive loader
    # would bypass the engine-backed model's ``load_state_dict`` and find no tensors to load into
    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        if not hasattr(self.model, "core"):
            return super().state_dict(*args, destination=destination, prefix=prefix, keep_vars=keep_vars)
        out = destination if destination is not None else {}
        for k, v in self.model.state_dict().items():
            out[prefix + "model." + k] = v
        return out

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        if not hasattr(self.model, "core"):
            return super().load_state_dict(state_dict, strict=strict, assign=assign)
        inner = {k[len("model."):]: v for k, v in state_dict.items() if k.startswith("model.")}
        res = self.model.load_state_dict(inner, strict=strict)
        unexpected = sorted(k for k in state_dict if not k.startswith("model."))
        if strict and unexpected:
            raise RuntimeError(f"unexpected keys in state_dict: {unexpected[:5]}")
        missing = ["model." + k for k in getattr(res, "missing_keys", [])]
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def _current_lr(self) -> float:
        """The learning rate Lightning's (possibly scheduled) optimizer holds right now; the factory's without a Trainer."""
        try:
            opt = self.optimizers()
        except Exception:  # noqa: BLE001 - no trainer attached (direct use, tests)
            opt = None
        if isinstance(opt, (list, tuple)):
            opt = opt[0] if opt else None
        if opt is not None and getattr(opt, "param_groups", None):
            return float(opt.param_groups[0]["lr"])
        return float(self._optimizer_factory.learning_rate)

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
            lr = self._current_lr()
            rw = spec.row_weights(batch["feature_tensors"], tm) if hasattr(spec, "row_weights") else None
            loss = core.fused_step(batch["feature_tensors"][core.item_feature], batch["padding_mask"], lab, tm,
                                   lr=lr, negatives=neg, row_weights=rw)  # all_reduce="auto": DDP gradient exchange inside
            self.log("learning_rate", lr, on_step=True, on_epoch=True, prog_bar=True, sync_dist=True)
        else:
            loss = self(batch)["loss"]
        self.log("train_loss", loss, on_step=True, on_epoch=True, prog_bar=True, sync_dist=True)
        return loss

    def on_train_epoch_end(self):
        # manual optimisation: Lightning does not step lr schedulers by itself (default interval of the reference's
        # factories: once per epoch, replay/nn/lightning/scheduler.py)
        if self.fused_optimizer and self._lr_scheduler_factory is not None:
            try:
                sch = self.lr_schedulers()
            except Exception:  # noqa: BLE001
                sch = None
            for s_ in (sch if isinstance(sch, (list, tuple)) else [sch]):
                if s_ is not None:
                    s_.step()

    def _inference(self, batch: dict):
        self.model.eval()
        if hasattr(self.model, "core"):
            return LazyInferenceOutput(lambda: self(batch))
        return self(batch)

    def predict_step(self, batch: dict, batch_idx: int = 0, dataloader_idx: int = 0):
        return self._inference(batch)

    def validation_step(self, batch: dict, batch_idx: int = 0, dataloader_idx: int = 0):
        return self._inference(batch)

    def test_step(self, batch: dict, batch_idx: int = 0, dataloader_idx: int = 0):
        return self._inference(batch)

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
