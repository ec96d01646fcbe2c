"""Mirror of ``replay.models.nn.sequential`` (legacy Lightning modules) for the hot path."""
from .callbacks import (PandasPredictionCallback, PolarsPredictionCallback, QueryEmbeddingsPredictionCallback,  # noqa: F401
                        SparkPredictionCallback, TorchPredictionCallback, ValidationMetricsCallback)
from .postprocessors import RemoveSeenItems  # noqa: F401
from .sasrec import SasRec, SasRecModel  # noqa: F401
from .bert4rec import Bert4Rec, Bert4RecModel, shift_features, uniform_masker  # noqa: F401
