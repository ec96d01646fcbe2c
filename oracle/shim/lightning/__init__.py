"""Stub `lightning` (not installed here): just enough surface for the reference modules to import (test infra only)."""
import torch


class _HParams(dict):
    __getattr__ = dict.get


class LightningModule(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self._hparams = _HParams()
        self.trainer = None

    def save_hyperparameters(self, *a, **k):
        pass

    @property
    def hparams(self):
        return self._hparams

    def log(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass


class LightningDataModule:
    def __init__(self, *a, **k):
        pass


class Callback:
    pass


class Trainer:
    def __init__(self, *a, **k):
        pass


from . import pytorch  # noqa: E402,F401
