from .sasrec import SasRec  # noqa: F401
