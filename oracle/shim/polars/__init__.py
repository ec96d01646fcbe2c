"""Stub `polars` so that the reference's pure-torch classes import in this container (test infrastructure only).

polars is not installed here and there is no network; replay/utils/types.py:6 imports it unconditionally.
Any attribute resolves to a dummy class so that type annotations / isinstance checks evaluate.
"""


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        return _Dummy()


class DataFrame(_Dummy):
    pass


class LazyFrame(_Dummy):
    pass


class Series(_Dummy):
    pass


class Expr(_Dummy):
    pass


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return type(name, (_Dummy,), {})
