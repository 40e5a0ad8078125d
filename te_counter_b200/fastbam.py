"""
ctypes binding of libtecbam.so (include/tecbam.h): BAM file -> the structure-of-arrays batches of
te_counter_b200/reads.py, decoded by a pool of host threads instead of a Python loop over pysam
records (reference read loops te_count/te_count.py:65-98, :190-214, :351-438).

`NativeBam.fill_bulk` / `.fill_sc` are drop-ins for `reads.fill_bulk` / `reads.fill_sc`: same
arrays, same sentinel values, same exceptions for the inputs on which the reference raises.
tests/test_fastbam.py holds the two against each other record for record.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TEC_BAM_LIB") or os.path.join(HERE, "libtecbam.so")

E_IO, E_FORMAT, E_NOT_BGZF, E_ARG = -1, -2, -3, -4
E_NO_BARCODE_TAG, E_NO_UMI_TAG, E_UMI, E_END_NONE, E_CHROM_NAME, E_REF_NONE = -10, -11, -12, -13, -14, -15
CHROM_SC_BAD = 0xFFFD

_vp, _i64, _i32, _int = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int
SIGNATURES = {
    "tbam_abi_version": (_int, []),
    "tbam_strerror": (ctypes.c_char_p, [_int]),
    "tbam_open": (_int, [ctypes.c_char_p, _int, ctypes.POINTER(_vp)]),
    "tbam_close": (None, [_vp]),
    "tbam_last_error": (ctypes.c_char_p, [_vp]),
    "tbam_n_references": (_int, [_vp]),
    "tbam_reference_name": (ctypes.c_char_p, [_vp, _int]),
    "tbam_set_chrom_map": (_int, [_vp, _vp, _vp, _i32, _i32]),
    "tbam_set_whitelist": (_int, [_vp, ctypes.c_char_p, _vp, _i32]),
    "tbam_next_bulk": (_int, [_vp, _int, _int, _i64] + [_vp] * 5 + [ctypes.POINTER(_i64), ctypes.POINTER(_int)]),
    "tbam_next_sc": (_int, [_vp, _int, _i64] + [_vp] * 7 + [ctypes.POINTER(_i64), ctypes.POINTER(_int)]),
    "tbam_counter": (_i64, [_vp, _int]),
    "tbam_inflate_raw": (_int, [ctypes.c_char_p, _i64, _vp, _i64, _int]),
}

_lib = None


class NotBgzf(Exception):
    """The file is not block-compressed BAM (SAM text, plain gzip): use another reader."""


def available():
    return os.path.exists(LIB_PATH)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s not built: run `python -m te_counter_b200.build`" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.tbam_abi_version() != 1:
            raise RuntimeError("libtecbam.so: ABI version mismatch")
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(_vp)


class NativeBam:
    def __init__(self, filename, threads=None):
        lib = load()
        if threads is None:
            threads = int(os.environ.get("TEC_BAM_THREADS", "0"))
        self._h = _vp()
        rc = lib.tbam_open(os.fsencode(filename), threads, ctypes.byref(self._h))
        if rc == E_NOT_BGZF:
            raise NotBgzf(filename)
        if rc == E_IO:
            raise OSError("cannot open %s" % filename)
        if rc:
            raise ValueError("%s: %s" % (filename, lib.tbam_strerror(rc).decode()))
        self.filename = filename
        self.references = [lib.tbam_reference_name(self._h, i).decode("ascii")
                           for i in range(lib.tbam_n_references(self._h))]

    def bind(self, chrom_map, whitelist=None):
        """Hands over what the reference's loop looks up per record: the chromosome key of every
        reference sequence (reads.ChromMap) and, for --sc, the sorted whitelist."""
        lib = load()
        bulk = np.array([chrom_map.bulk_id(n) for n in self.references], dtype=np.uint16)
        sc = np.empty(len(self.references), dtype=np.uint16)
        for i, n in enumerate(self.references):
            try:
                sc[i] = chrom_map.sc_id(n)
            except ValueError:                      # ':' in the name: raised when a record gets there
                sc[i] = CHROM_SC_BAD
        self._check(lib.tbam_set_chrom_map(self._h, _p(bulk), _p(sc), len(self.references), chrom_map.n_index))
        if whitelist is not None:
            enc = [b.encode("utf-8") for b in whitelist.id_to_barcode]
            off = np.zeros(len(enc) + 1, dtype=np.int64)
            np.cumsum([len(b) for b in enc], out=off[1:])
            self._check(lib.tbam_set_whitelist(self._h, b"".join(enc), _p(off), len(enc)))

    def _check(self, rc, bulk=False):
        if rc == 0:
            return
        msg = load().tbam_last_error(self._h).decode()
        if rc in (E_NO_BARCODE_TAG, E_NO_UMI_TAG):
            raise AssertionError(msg.split(" (record")[0])                  # te_count.py:409, :426
        if rc == E_END_NONE:
            raise TypeError("unsupported operand type(s) for +: 'NoneType' and 'int'" if bulk else msg)
        if rc == E_REF_NONE:
            raise AttributeError("'NoneType' object has no attribute 'replace'")    # te_count.py:431
        if rc == E_FORMAT and "truncated" in msg:
            raise EOFError("%s: %s" % (self.filename, msg))
        raise ValueError("%s: %s" % (self.filename, msg))

    def fill_bulk(self, batch, paired, qual):
        n, more = _i64(0), _int(0)
        rc = load().tbam_next_bulk(self._h, int(bool(paired)), int(qual), batch.capacity, _p(batch.start), _p(batch.end),
                                   _p(batch.chrom), _p(batch.mapq), _p(batch.flag), ctypes.byref(n), ctypes.byref(more))
        self._check(rc, bulk=True)
        batch.n = n.value
        return bool(more.value)

    def fill_sc(self, batch, qual):
        n, more = _i64(0), _int(0)
        rc = load().tbam_next_sc(self._h, int(qual), batch.capacity, _p(batch.start), _p(batch.end), _p(batch.chrom),
                                 _p(batch.mapq), _p(batch.flag), _p(batch.cell), _p(batch.umi),
                                 ctypes.byref(n), ctypes.byref(more))
        self._check(rc)
        batch.n = n.value
        return bool(more.value)

    def counters(self):
        lib = load()
        names = ("records", "compressed_bytes", "uncompressed_bytes", "threads", "ns_in_next")
        return {k: int(lib.tbam_counter(self._h, i)) for i, k in enumerate(names)}

    def close(self):
        if self._h:
            load().tbam_close(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
