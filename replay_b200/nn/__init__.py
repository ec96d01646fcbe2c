"""Mirror of ``replay.nn`` for the sequential hot path (new block-based API)."""
from .sequential import SasRec  # noqa: F401
