"""Mirror of ``replay.nn.lightning``: universal LightningModule wrapper, optimizer factory, top-items callbacks and the
seen-items postprocessor (module.py:13-123, optimizer.py:24-60, callback/predictions_callback.py:29-163,
postprocessor/seen_items.py:8-83)."""
from .callback import (HiddenStatesCallback, PandasTopItemsCallback, PolarsTopItemsCallback, SparkTopItemsCallback,  # noqa: F401
                       TopItemsCallbackBase, TorchTopItemsCallback)
from .metrics import ComputeMetricsCallback, RankingMetrics, TorchMetricsBuilder  # noqa: F401
from .module import LightningModule, OptimizerFactory  # noqa: F401
from .postprocessor import SeenItemsFilter  # noqa: F401
