"""TEST INFRASTRUCTURE -- ctypes loader of oracle/libteoracle.so (C restatement of the bulk loop,
oracle/te_oracle_c.c).  Same rules as oracle/te_oracle.py: only tests, smoke() and bench.py's
parity / cpu legs may use it, as the checker."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libteoracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "te_oracle_c.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        lib = ctypes.CDLL(LIB)
        vp = ctypes.c_void_p
        lib.teo_index_build.restype = vp
        lib.teo_index_build.argtypes = [ctypes.c_int64, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.teo_index_free.argtypes = [vp]
        lib.teo_bulk_count.restype = ctypes.c_int
        lib.teo_bulk_count.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int64, vp, vp, vp, vp, vp, vp, vp, ctypes.c_int]
        _lib = lib
    return _lib


class Index:
    """Arrays in linearData order, as oracle/te_oracle.Index."""

    def __init__(self, chrom_id, L, R, ensg_id, type_code, n_chrom, n_ensg, bucket_size=10000):
        self._keep = [np.ascontiguousarray(chrom_id, np.int32), np.ascontiguousarray(L, np.int32),
                      np.ascontiguousarray(R, np.int32), np.ascontiguousarray(ensg_id, np.int32),
                      np.ascontiguousarray(type_code, np.uint8)]
        self.n_ensg = int(n_ensg)
        self._lib = load()
        self._h = self._lib.teo_index_build(len(self._keep[1]), *[a.ctypes.data for a in self._keep],
                                            int(n_chrom), int(n_ensg), int(bucket_size))

    def __del__(self):
        try:
            self._lib.teo_index_free(self._h)
        except Exception:
            pass


def bulk_count(idx, paired, qual, start, end, chrom, mapq, flag, threads=0):
    """(counts int64[n_ensg], stats int64[7] in include/tecount.h order).  Arrays may be numpy arrays
    or raw host pointers (ints) with n given by len(start) / the `n` of a (ptr, n) tuple."""
    lib = load()
    cols = []
    n = None
    for a, dt in zip((start, end, chrom, mapq, flag), (np.int32, np.int32, np.uint16, np.uint8, np.uint8)):
        a = np.ascontiguousarray(a, dtype=dt)
        n = len(a) if n is None else n
        cols.append(a)
    counts = np.zeros(idx.n_ensg, np.int64)
    stats = np.zeros(7, np.int64)
    lib.teo_bulk_count(idx._h, 1 if paired else 0, int(qual), int(n), *[c.ctypes.data for c in cols],
                       counts.ctypes.data, stats.ctypes.data, int(threads or os.cpu_count() or 1))
    return counts, stats
