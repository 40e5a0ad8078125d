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
        lib.teo_sc_count.restype = vp
        lib.teo_sc_count.argtypes = [ctypes.c_int64, vp, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                     ctypes.c_int64, vp, vp, vp, vp, vp, vp, vp]
        lib.teo_sc_n_triples.restype = ctypes.c_int64
        lib.teo_sc_n_triples.argtypes = [vp]
        lib.teo_sc_n_hit.restype = ctypes.c_int64
        lib.teo_sc_n_hit.argtypes = [vp]
        lib.teo_sc_fetch.argtypes = [vp] * 7
        lib.teo_sc_free.argtypes = [vp]
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


def sc_count(feat, n_chrom, bucket_size, qual, strand, bundle_keys, maxcells, pad, start, end, chrom, mapq, flag, cell, umi):
    """C++ restatement of te_oracle.sc_count (oracle/te_oracle_sc.cpp).  feat = (chrom_id, L, R, ensg_id,
    type_code, strand_code) in linearData order.  Returns the same dict as the Python oracle
    (ReferenceCrash cases are reported in stats['crash_strand'] / stats['zero_division'])."""
    lib = load()
    f = [np.ascontiguousarray(a, dt) for a, dt in zip(feat, (np.int32, np.int32, np.int32, np.int32, np.uint8, np.uint8))]
    cols = [np.ascontiguousarray(a, dt) for a, dt in zip((start, end, chrom, mapq, flag, cell, umi),
                                                         (np.int32, np.int32, np.uint16, np.uint8, np.uint8, np.uint32, np.uint64))]
    h = lib.teo_sc_count(len(f[1]), *[a.ctypes.data for a in f], int(n_chrom), int(bucket_size), int(qual), 1 if strand else 0,
                         int(bundle_keys), int(maxcells), int(pad), len(cols[0]), *[c.ctypes.data for c in cols])
    nt, nh = lib.teo_sc_n_triples(h), lib.teo_sc_n_hit(h)
    ensg = np.zeros(nt, np.int32); tcell = np.zeros(nt, np.uint32); cnt = np.zeros(nt, np.int64)
    hcell = np.zeros(nh, np.uint32); hcnt = np.zeros(nh, np.int64); st = np.zeros(12, np.int64)
    lib.teo_sc_fetch(h, *[a.ctypes.data for a in (ensg, tcell, cnt, hcell, hcnt, st)])
    lib.teo_sc_free(h)
    keys = ("total_reads", "invalid_barcode", "already_seen", "lowq", "qcfail", "valid", "assigned", "raw_barcodes",
            "n_bundles", "crash_strand", "zero_division")
    return {"triples": {(int(e), int(c)): int(v) for e, c, v in zip(ensg, tcell, cnt)},
            "triples_arrays": (ensg, tcell, cnt),
            "cell_hits": list(zip(hcell.tolist(), hcnt.tolist())),
            "stats": {k: int(st[i]) for i, k in enumerate(keys)}}
