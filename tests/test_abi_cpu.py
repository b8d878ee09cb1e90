"""CPU-only checks of the C-ABI boundary: the library loads, exports every symbol that
include/zc_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import REPO
from zeroclone_b200 import _ffi
from zeroclone_b200.build import build


@pytest.fixture(scope="module")
def L():
    build()
    return _ffi.lib()


def declared_symbols():
    src = open(os.path.join(REPO, "include", "zc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zc_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(L):
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/zc_b200.h but not exported"


def test_abi_version_and_struct_sizes(L):
    assert L.zc_abi_version() == 2
    assert C.sizeof(_ffi.C4State) == 24 and C.sizeof(_ffi.ChessState) == 72 and C.sizeof(_ffi.RootResult) == 48


def test_move_order_roundtrip_and_validation(L):
    from oracle import zc_oracle as zo
    t = zo.python_c4_order()
    _ffi.check(L.zc_c4_set_move_order(t.ctypes.data_as(C.c_void_p)))
    back = np.zeros((128, 8), dtype=np.uint8)
    _ffi.check(L.zc_c4_get_move_order(back.ctypes.data_as(C.c_void_p)))
    assert (back == t).all()
    bad = t.copy()
    bad[127, 0] = bad[127, 1]
    assert L.zc_c4_set_move_order(bad.ctypes.data_as(C.c_void_p)) == _ffi.ZC_EINVAL


def test_no_cpu_fallback_without_gpu(L):
    if L.zc_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = L.zc_search_create(_ffi.GAME_C4, 0, 4, 32, 0, C.byref(h))
    assert rc == _ffi.ZC_ENODEVICE
    assert b"no CPU path" in L.zc_last_error()


def test_tower_refuses_without_gpu_and_checks_arguments(L):
    w = np.zeros(128 * 2 * 9 + 16 * 128 * 128 * 9, dtype=np.float32)
    b = np.zeros(17 * 128, dtype=np.float32)
    hw = np.zeros(128, dtype=np.float32)
    h = C.c_void_p()
    args = (w.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), hw.ctypes.data_as(C.c_void_p), C.c_float(0.0), C.byref(h))
    f16 = _ffi.PLANE_F16
    assert L.zc_tower_create(7, 0, 8, f16, *args) == _ffi.ZC_EINVAL            # unknown game
    assert L.zc_tower_create(_ffi.GAME_C4, 0, 9, f16, *args) == _ffi.ZC_EINVAL   # deeper than the reference tower
    assert L.zc_tower_create(_ffi.GAME_C4, 0, 8, _ffi.PLANE_F32, *args) == _ffi.ZC_EINVAL   # tensor-core operands are 16-bit
    assert L.zc_tower_update_weights(None, *args[:4], None) == _ffi.ZC_EINVAL
    assert L.zc_tower_fault(None) == 0
    if L.zc_device_count() > 0:
        pytest.skip("a GPU is present")
    assert L.zc_tower_create(_ffi.GAME_C4, 0, 8, f16, *args) == _ffi.ZC_ENODEVICE
    assert b"no CPU path" in L.zc_last_error()
    assert L.zc_tower_forward(None, None, 4, None, None) == _ffi.ZC_EINVAL
