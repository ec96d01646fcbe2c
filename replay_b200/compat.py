"""Lightning compatibility: subclass ``lightning.LightningModule`` / ``Callback`` when Lightning is importable, otherwise
minimal stand-ins exposing the hooks the reference's modules use (training_step / predict_step / configure_optimizers /
log / save_hyperparameters), so the mirrors work - and are testable - where Lightning is not installed."""
from __future__ import annotations

import torch

try:  # pragma: no cover - depends on the environment
    import lightning as _L

    LightningModuleBase = _L.LightningModule
    CallbackBase = _L.Callback
    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    HAVE_LIGHTNING = False

    class _HParams(dict):
        __getattr__ = dict.get

    class LightningModuleBase(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self._hparams = _HParams()
            self.trainer = None
            self.automatic_optimization = True
            self.logged = {}

        def save_hyperparameters(self, *names, ignore=None, **k):
            import inspect

            frame = inspect.currentframe().f_back
            loc = frame.f_locals
            ignore = set(ignore or [])
            for key, val in loc.items():
                if key in ("self", "__class__") or key in ignore:
                    continue
                self._hparams[key] = val

        @property
        def hparams(self):
            return self._hparams

        def log(self, name, value, *a, **k):
            self.logged[name] = value

        def log_dict(self, d, *a, **k):
            self.logged.update(d)

    class CallbackBase:
        pass
