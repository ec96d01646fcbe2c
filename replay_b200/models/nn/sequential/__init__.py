"""Mirror of ``replay.models.nn.sequential`` (legacy Lightning modules) for the hot path."""
from .callbacks import PandasPredictionCallback, TorchPredictionCallback  # noqa: F401
from .postprocessors import RemoveSeenItems  # noqa: F401
from .sasrec import SasRec, SasRecModel  # noqa: F401
from .bert4rec import Bert4Rec, Bert4RecModel, shift_features, uniform_masker  # noqa: F401
