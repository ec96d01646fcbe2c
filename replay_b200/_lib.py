"""ctypes binding of librp_b200.so (C ABI: include/rp_b200.h).  Fails loudly when the library is missing - there is no
CPU or PyTorch fallback for the kernels."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librp_b200.so")

_lib = None


class RpError(RuntimeError):
    pass


_ERR = {-1: "RP_EINVAL (null pointer / unsupported flag)", -2: "RP_ESHAPE (unsupported size)",
        -3: "RP_EALIGN (pointer or pitch not 16-byte aligned)", -4: "RP_EDRIVER (driver entry point / tensor map)",
        -5: "RP_EWORKSPACE (workspace too small)"}


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc < 0:
        raise RpError(f"{what}: {_ERR.get(rc, rc)}")
    raise RpError(f"{what}: cudaError {rc}")


def _sig(fn, restype, argtypes):
    fn.restype = restype
    fn.argtypes = argtypes


def lib():
    """Load (once) and return the ctypes handle.  Raises if the extension has not been built (python -m replay_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RpError(
            f"{LIB_PATH} is missing: build it with `python -m replay_b200.build` (nvcc, sm_100a). "
            "replay_b200 has no CPU fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    P = c_void_p
    _sig(L.rp_version, c_char_p, [])
    _sig(L.rp_selftest_umma, c_int, [c_int, P, P, P, P])
    _sig(L.rp_seen_prepare, c_int, [P, c_int, c_int, c_int, P, P, P])
    _sig(L.rp_score_topk_workspace, c_size_t, [c_int, c_int, c_int, c_int])
    _sig(L.rp_score_topk, c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, P, c_size_t, P])
    _sig(L.rp_ce_head_workspace, c_size_t, [c_int, c_int, c_int])
    _sig(L.rp_ce_head_fwd, c_int, [P, P, P, P, c_int, c_int, c_int, P, P, P, P, c_size_t, P])
    _sig(L.rp_ce_head_bwd, c_int, [P, P, P, P, c_int, c_int, c_int, P, P, P, P, P])
    for name, restype, argtypes in _EXTRA_SIGS:
        _sig(getattr(L, name), restype, argtypes)
    _lib = L
    return L


# filled in by replay_b200.ops as kernels are added (keeps one place per kernel family)
_EXTRA_SIGS: list = []

__all__ = ["lib", "check", "RpError", "LIB_PATH", "c_float", "c_int", "c_int32", "c_int64", "c_size_t", "c_void_p"]
