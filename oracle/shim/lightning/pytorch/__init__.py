from .. import Callback, LightningDataModule, LightningModule, Trainer  # noqa: F401
from . import callbacks, trainer, utilities  # noqa: F401
