"""libtecount.so loads on a CPU-only box and exports every symbol include/tecount.h declares
(no compute call is made here)."""
import ctypes
import os
import re

from te_counter_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "tecount.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tec_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    build.build()
    names = declared_symbols()
    assert len(names) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libtecount.so does not export %s" % n
        assert n in _lib.SIGNATURES, "ctypes binding lacks %s" % n
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_strerror():
    lib = _lib.load_library()
    assert lib.tec_abi_version() == 1
    assert lib.tec_strerror(0) == b"ok"
    assert b"limit" in lib.tec_strerror(-5)
