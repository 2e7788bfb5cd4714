"""Loader of the CUDA library (``libmadigan_b200.so``) through ctypes.

There is no CPU fallback: if the library is missing this module raises, and every
environment operation goes through the C-ABI of ``include/madigan_b200.h``.
"""
import ctypes as C
import os

from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
_VARIANT = os.environ.get("MDG_LIB_VARIANT", "")  # profiling only, see build.py
LIB_PATH = os.path.join(_HERE, f"libmadigan_b200.{_VARIANT}.so" if _VARIANT else "libmadigan_b200.so")
_lib = None


class MadiganCudaError(RuntimeError):
    pass


def lib():
    """The bound CUDA library; raises ImportError (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built. "
                "Run `python -m madigan_b200.build` (needs nvcc); there is no CPU fallback.")
        L = A.bind(C.CDLL(LIB_PATH))
        v = L.mdg_abi_version()
        if v != A.MDG_ABI_VERSION:
            raise ImportError(f"libmadigan_b200.so has ABI {v}, python expects {A.MDG_ABI_VERSION}: rebuild")
        _lib = L
    return _lib


def check(rc):
    """Map MDG_E_* to the exception classes the reference's pybind layer raises
    (reference: DataTypes.h:36-46 -> RuntimeError; std::out_of_range -> IndexError)."""
    if rc == A.MDG_OK:
        return
    msg = lib().mdg_last_error().decode(errors="replace")
    if rc == A.MDG_E_INVALID:
        if "out of range" in msg:
            raise IndexError(msg)
        raise ValueError(msg)
    if rc == A.MDG_E_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise MadiganCudaError(msg)
