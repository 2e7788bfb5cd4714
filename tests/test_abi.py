"""CPU checks of the C-ABI boundary: the library loads, exports every symbol the header declares,
and its structs have the sizes the ctypes mirror assumes.  No kernel is launched."""
import ctypes as C
import os
import re

import pytest

from madigan_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "madigan_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mdg_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def cdll():
    from madigan_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from madigan_b200.build import build
        build()
    return _lib.lib()


def test_header_and_python_mirror_agree():
    assert declared_symbols() == sorted(A.SYMBOLS)


def test_library_exports_every_declared_symbol(cdll):
    for name in declared_symbols():
        assert hasattr(cdll, name), f"{name} missing from libmadigan_b200.so"
    assert cdll.mdg_abi_version() == A.MDG_ABI_VERSION


def test_struct_sizes_match(cdll):
    for i, st in enumerate(A.STRUCTS):
        assert cdll.mdg_sizeof(i) == C.sizeof(st), st.__name__
    assert cdll.mdg_sizeof(99) == -1


def test_argument_validation_without_gpu(cdll):
    """Launchers validate arguments before touching CUDA; errors map to the reference's exception types."""
    from madigan_b200._lib import check
    from madigan_b200.environments.data_source import make_params
    P, _ = make_params("OU")
    L = A.MdgLaunch(n_envs=4, window=8, head=9)  # head outside the ring
    S, IO = A.MdgState(), A.MdgStepIO()
    S.folds = 1  # non-null dummies; never dereferenced because validation fails first
    with pytest.raises(ValueError):
        check(cdll.mdg_step(C.byref(P), None, C.byref(S), C.byref(IO), C.byref(L)))
    L = A.MdgLaunch(n_envs=4, window=8, head=0, mode=A.MODE_SINGLE, asset_idx=7)
    IO.units = 1  # non-null dummy; never dereferenced because validation fails first
    with pytest.raises(IndexError):  # std::out_of_range -> IndexError in the reference (envTest.py:245-248)
        check(cdll.mdg_step(C.byref(P), None, C.byref(S), C.byref(IO), C.byref(L)))
    P.n_assets = 99
    with pytest.raises(NotImplementedError):
        check(cdll.mdg_step(C.byref(P), None, C.byref(S), C.byref(IO), C.byref(L)))
    # zero envs is a no-op, not an error
    P.n_assets = 4
    L = A.MdgLaunch(n_envs=0, window=8, head=0, mode=A.MODE_HOLD)
    assert cdll.mdg_step(C.byref(P), None, C.byref(S), C.byref(IO), C.byref(L)) == 0
