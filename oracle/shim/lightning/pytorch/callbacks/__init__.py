from ... import Callback  # noqa: F401
