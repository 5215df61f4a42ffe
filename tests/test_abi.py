"""CPU checks: the C-ABI library builds/loads and exports every symbol include/sat_b200.h declares,
and the ctypes struct mirrors agree with the compiled sizes (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sat_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sat_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    import sat_b200
    from sat_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    L = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 8
    for s in syms:
        assert hasattr(L, s), "libsat_b200.so does not export %s" % s
    for s in _lib.EXPORTS:
        assert s in syms, "%s bound in _lib.py but not declared in the header" % s
    assert L.sat_version() >= 1


def test_struct_mirrors_match():
    from sat_b200 import _lib
    L = _lib.lib()
    for i, st in enumerate((_lib.SatDims, _lib.SatWeights, _lib.SatTrainBuffers, _lib.SatDecodeBuffers, _lib.SatMasterWeights)):
        assert L.sat_abi_sizeof(i) == ctypes.sizeof(st)


def test_bad_arguments_fail_loudly():
    from sat_b200 import _lib
    L = _lib.lib()
    d = _lib.SatDims()
    d.B, d.Bi, d.ncap, d.L, d.D, d.A, d.E, d.H, d.V, d.T, d.dtype = 2, 2, 1, 4, 12, 8, 8, 8, 8, 1, 0   # D not a multiple of 8
    w = _lib.SatWeights()
    rc = L.sat_prepare_images(ctypes.byref(d), ctypes.byref(w), None, None, None, None, None, None, None, None)
    assert rc < 0 and b"multiples of 8" in L.sat_last_error()


def test_no_cpu_fallback():
    import torch
    from sat_b200 import _lib
    with pytest.raises(AssertionError):
        _lib.ptr(torch.zeros(4))
