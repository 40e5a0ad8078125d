"""
Reader / flattener for the `.glb` annotation index that te_genome builds.

The index is a pickle of a `te_count.miniglbase.genelist.genelist` whose rows hold a
`te_count.miniglbase.location.location` (reference: miniglbase/base_genelist.py:302-306 `save`,
miniglbase/utils.py:21-59 `glload`, te_count.py:31-35 `load_genome`).  This module reads that
format WITHOUT importing the reference package: a restricted unpickler maps the two class paths
onto attribute bags, then the rows are flattened to the arrays the CUDA library consumes:

    per feature, in linearData order : chrom_id, L, R, ensg_id, type_code, strand_code
    per chromosome, sorted by (L, R) : the same columns + offsets          (device layout)
    names                            : sorted(set(ensg))  == TSV row / column order (te_count.py:35)

The reference's candidate set is the 10 kb bucket hash stored in the pickle
(`genelist.buckets`, miniglbase/genelist.py:367-380).  The kernels use its closed form
(L//bs <= b <= R//bs), so `verify_buckets` checks the stored hash against that closed form and the
load fails loudly if they differ (e.g. an index written with another bucket_size).
"""
import io
import os
import pickle

import numpy as np

BUCKET_SIZE = 10000                      # miniglbase/config.py:36

T_OTHER, T_GENE, T_TE, T_SNRNA, T_ENH = 0, 1, 2, 3, 4
_TYPE_CODES = {"protein_coding": T_GENE, "lincRNA": T_GENE, "lncRNA": T_GENE,   # te_count.py:134
               "TE": T_TE,                                                         # :139
               "snRNA": T_SNRNA,                                                   # :142
               "enhancer": T_ENH}                                                  # :145
STRAND_MISSING = 255
MAX_ENSG = 1 << 24                       # ensg id field width in the packed device word


class _Bag:
    """Attribute bag standing in for genelist / location instances while unpickling.  No
    __setstate__: the C unpickler then restores the instance dict (and a (dict, slots) pair) itself,
    which is ~6 us per feature cheaper than a Python-level call."""


class _Genelist(_Bag):
    pass


class _Location(_Bag):
    pass


_CLASS_MAP = {
    ("te_count.miniglbase.genelist", "genelist"): _Genelist,
    ("te_count.miniglbase.location", "location"): _Location,
    # the same classes when the index was written by a stand-alone glbase3 / miniglbase
    ("miniglbase.genelist", "genelist"): _Genelist,
    ("miniglbase.location", "location"): _Location,
    ("glbase3.genelist", "genelist"): _Genelist,
    ("glbase3.location", "location"): _Location,
}
_SAFE_BUILTINS = {"set", "frozenset", "list", "dict", "tuple", "int", "float", "str", "bool",
                  "bytes", "complex", "slice", "range"}


class _GlbUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) in _CLASS_MAP:
            return _CLASS_MAP[(module, name)]
        if module == "builtins" and name in _SAFE_BUILTINS:
            return getattr(__import__("builtins"), name)
        if module == "collections" and name in ("OrderedDict", "defaultdict"):
            import collections
            return getattr(collections, name)
        raise pickle.UnpicklingError("refusing to unpickle %s.%s from a .glb index" % (module, name))


class GlbIndex:
    """Flattened index.  Attributes (numpy, linearData order):
    chrom_id i32, L i32, R i32, ensg_id i32, type_code u8, strand_code u8;
    names (sorted unique ensg), chrom_keys (normalised chromosome strings, id order),
    strand_strings (code -> string)."""

    def __init__(self, chrom_keys, chrom_id, L, R, ensg_id, type_code, strand_code, names,
                 strand_strings=("+", "-"), bucket_size=BUCKET_SIZE, type_strings=None):
        self.chrom_keys = list(chrom_keys)
        self.chrom_lookup = {k: i for i, k in enumerate(self.chrom_keys)}
        self.chrom_id = np.ascontiguousarray(chrom_id, dtype=np.int32)
        self.L = np.ascontiguousarray(L, dtype=np.int32)
        self.R = np.ascontiguousarray(R, dtype=np.int32)
        self.ensg_id = np.ascontiguousarray(ensg_id, dtype=np.int32)
        self.type_code = np.ascontiguousarray(type_code, dtype=np.uint8)
        self.strand_code = np.ascontiguousarray(strand_code, dtype=np.uint8)
        self.names = list(names)
        self.strand_strings = list(strand_strings)
        self.bucket_size = int(bucket_size)
        self.type_strings = type_strings
        n = len(self.L)
        assert all(len(a) == n for a in (self.chrom_id, self.R, self.ensg_id, self.type_code,
                                         self.strand_code))
        if len(self.names) > MAX_ENSG:
            raise ValueError("index has %d distinct ensg names; the device word holds %d"
                             % (len(self.names), MAX_ENSG))
        self._sorted = None

    @property
    def n_features(self):
        return len(self.L)

    @property
    def n_ensg(self):
        return len(self.names)

    @property
    def n_chrom(self):
        return len(self.chrom_keys)

    def sorted_layout(self):
        """Device layout: features grouped by chromosome id and sorted by (L, R, original index).
        Returns dict(chrom_off i64[n_chrom+1], L, R, ensg_id, type_code, strand_code, order)."""
        if self._sorted is None:
            n = self.n_features
            order = np.lexsort((np.arange(n), self.R, self.L, self.chrom_id))
            cid = self.chrom_id[order]
            off = np.zeros(self.n_chrom + 1, dtype=np.int64)
            if n:
                off[1:] = np.cumsum(np.bincount(cid, minlength=self.n_chrom))
            self._sorted = {
                "chrom_off": off,
                "L": np.ascontiguousarray(self.L[order]),
                "R": np.ascontiguousarray(self.R[order]),
                "ensg_id": np.ascontiguousarray(self.ensg_id[order]),
                "type_code": np.ascontiguousarray(self.type_code[order]),
                "strand_code": np.ascontiguousarray(self.strand_code[order]),
                "order": order,
            }
        return self._sorted


def _read_pickle(path):
    import gc
    was = gc.isenabled()
    gc.disable()                    # millions of small objects, none of them garbage: the collector only costs time here
    try:
        with open(os.path.realpath(path), "rb") as fh:
            return _GlbUnpickler(io.BufferedReader(fh)).load()
    finally:
        if was:
            gc.enable()


def verify_buckets(gl_buckets, chrom_str, L, R, bs=BUCKET_SIZE):
    """Check the pickled bucket hash == {b*bs: [n : L[n]//bs <= b <= R[n]//bs]} per chromosome
    (miniglbase/genelist.py:367-380).  Raises ValueError when it is anything else.

    One flat pass: every (chromosome, bucket, feature id) entry of the hash goes into three arrays
    (C-level iteration over the dict values), then the membership rule, the chromosome key, the
    absence of duplicates and the total are checked with array operations -- a genome-scale index
    has ~300 k buckets, a Python loop over them cost more than the unpickling."""
    from itertools import chain
    n = len(L)
    L = np.asarray(L, dtype=np.int64)
    R = np.asarray(R, dtype=np.int64)
    expected_total = int(np.sum(R // bs - L // bs + 1)) if n else 0
    chrom_code = {}
    feat_chrom = np.fromiter((chrom_code.setdefault(c, len(chrom_code)) for c in chrom_str), dtype=np.int64, count=n)
    b_key, b_chrom, b_len = [], [], []
    flat = []
    for chrom, bdict in gl_buckets.items():
        code = chrom_code.get(chrom, -1)
        m = len(bdict)
        if not m:
            continue
        try:
            keys = np.fromiter(bdict.keys(), dtype=np.int64, count=m)
        except (TypeError, ValueError):
            raise ValueError("bucket keys of chromosome %r are not integers" % (chrom,))
        b_key.append(keys)
        b_chrom.append(np.full(m, code, dtype=np.int64))
        b_len.append(np.fromiter(map(len, bdict.values()), dtype=np.int64, count=m))
        flat.append(bdict.values())
    if not b_key:
        if expected_total:
            raise ValueError("bucket hash has 0 entries, closed form expects %d (bucket_size mismatch?)" % expected_total)
        return
    b_key, b_chrom, b_len = np.concatenate(b_key), np.concatenate(b_chrom), np.concatenate(b_len)
    total = int(b_len.sum())
    try:
        ids = np.fromiter(chain.from_iterable(chain.from_iterable(flat)), dtype=np.int64, count=total)
    except (TypeError, ValueError):
        raise ValueError("bucket hash holds something that is not a feature id")
    which = np.repeat(np.arange(len(b_key)), b_len)            # bucket of every entry
    if total and (ids.min() < 0 or ids.max() >= n):
        bad = which[np.flatnonzero((ids < 0) | (ids >= n))[0]]
        raise ValueError("bucket %s holds duplicate or out-of-range feature ids" % (int(b_key[bad]),))
    if (b_key % bs).any():
        raise ValueError("bucket %s is not the 10 kb bucket hash of its features" % (int(b_key[np.flatnonzero(b_key % bs)[0]]),))
    eb = b_key[which]
    ok = ((L[ids] // bs) * bs <= eb) & (eb <= (R[ids] // bs) * bs)
    if not ok.all():
        raise ValueError("bucket %s is not the 10 kb bucket hash of its features" % (int(eb[np.flatnonzero(~ok)[0]]),))
    wrong = feat_chrom[ids] != b_chrom[which]
    if wrong.any():
        raise ValueError("bucket chromosome key does not match its features (bucket %s)" % (int(eb[np.flatnonzero(wrong)[0]]),))
    pair = which * np.int64(max(n, 1)) + ids
    if len(np.unique(pair)) != total:
        raise ValueError("a bucket holds duplicate or out-of-range feature ids")
    if total != expected_total:
        raise ValueError("bucket hash has %d entries, closed form expects %d (bucket_size mismatch?)"
                         % (total, expected_total))


def from_rows(rows, buckets=None, verify=True):
    """rows: sequence of dicts with keys loc (chr/left/right mapping), type, ensg [, strand].
    Mirrors what load_genome() exposes: all_feature_names = sorted(set(ensg)) (te_count.py:35) and
    chromosome keys = the keys of genelist.buckets (first-appearance order of loc['chr']).

    Column-wise: every field is pulled out of the row dicts with C-level iteration (map +
    itemgetter) -- a genome-scale index has ~6 M rows and a Python statement per field per row
    costs more than unpickling them."""
    from itertools import repeat
    from operator import attrgetter, itemgetter
    rows = rows if isinstance(rows, list) else list(rows)
    n = len(rows)
    if any(map(contains_tss, rows)):
        raise ValueError("index rows carry 'tss_loc'; the reference would bucket on it "
                         "(miniglbase/genelist.py:345) -- unsupported")
    locs = list(map(itemgetter("loc"), rows))
    if n and not isinstance(locs[0], dict):
        try:
            locs = list(map(attrgetter("loc"), locs))           # location objects hold the dict in .loc
        except AttributeError:
            locs = [l.loc if hasattr(l, "loc") else l for l in locs]
    elif any(not isinstance(l, dict) for l in locs):
        locs = [l.loc if hasattr(l, "loc") else l for l in locs]
    chrom_str = list(map(itemgetter("chr"), locs))
    chrom_keys = list(dict.fromkeys(chrom_str))                  # first-appearance order
    chrom_lookup = {k: i for i, k in enumerate(chrom_keys)}
    chrom_id = np.fromiter(map(chrom_lookup.__getitem__, chrom_str), dtype=np.int32, count=n)
    L = np.fromiter(map(itemgetter("left"), locs), dtype=np.int64, count=n).astype(np.int32)
    R = np.fromiter(map(itemgetter("right"), locs), dtype=np.int64, count=n).astype(np.int32)
    type_code = np.fromiter(map(_TYPE_CODES.get, map(itemgetter("type"), rows), repeat(T_OTHER)), dtype=np.uint8, count=n)
    missing = object()
    strands = [r.get("strand", missing) for r in rows]
    strand_strings = ["+", "-"]
    for st in dict.fromkeys(strands):                            # distinct values in first-appearance order
        if st is not missing and st not in strand_strings:
            if len(strand_strings) >= 4:
                raise ValueError("more than 4 distinct strand strings in the index")
            strand_strings.append(st)
    strand_lookup = {st: i for i, st in enumerate(strand_strings)}
    strand_lookup[missing] = STRAND_MISSING
    strand_code = np.fromiter(map(strand_lookup.__getitem__, strands), dtype=np.uint8, count=n)
    ensg = list(map(itemgetter("ensg"), rows))
    names = sorted(set(ensg))
    name_id = {k: i for i, k in enumerate(names)}
    ensg_id = np.fromiter(map(name_id.__getitem__, ensg), dtype=np.int32, count=n)
    if buckets is not None and verify:
        verify_buckets(buckets, chrom_str, L, R)
        if list(buckets.keys()) != chrom_keys:
            raise ValueError("bucket chromosome keys differ from the feature rows")
    return GlbIndex(chrom_keys, chrom_id, L, R, ensg_id, type_code, strand_code, names, strand_strings)


def contains_tss(row):
    return "tss_loc" in row


CACHE_SUFFIX = ".tecidx.npz"
CACHE_VERSION = 1


def _fingerprint(path):
    """Size and BLAKE2b digest of the whole .glb: the sidecar is only used for exactly the file
    it was derived from (about 1 s per GB, against ~8.5 us per feature for the unpickle)."""
    import hashlib
    h = hashlib.blake2b(digest_size=20)
    with open(path, "rb") as fh:
        while True:
            b = fh.read(1 << 24)
            if not b:
                break
            h.update(b)
    return "%d:%s" % (os.path.getsize(path), h.hexdigest())


def _cache_enabled(cache):
    if cache is None:
        return os.environ.get("TEC_INDEX_CACHE", "1") not in ("0", "", "off", "no")
    return bool(cache)


def _load_sidecar(side, fingerprint):
    import json
    try:
        with np.load(side, allow_pickle=False) as z:
            meta = json.loads(bytes(z["meta"]).decode("utf-8"))
            if meta.get("version") != CACHE_VERSION or meta.get("fingerprint") != fingerprint:
                return None
            idx = GlbIndex(meta["chrom_keys"], z["chrom_id"], z["L"], z["R"], z["ensg_id"], z["type_code"],
                           z["strand_code"], meta["names"], meta["strand_strings"], bucket_size=meta["bucket_size"])
    except Exception:                                   # noqa: BLE001 -- unreadable, damaged or foreign file:
        return None                                     # whatever it is, the .glb is read instead
    return idx


def _write_sidecar(side, fingerprint, idx):
    import json
    meta = {"version": CACHE_VERSION, "fingerprint": fingerprint, "names": idx.names, "chrom_keys": idx.chrom_keys,
            "strand_strings": idx.strand_strings, "bucket_size": idx.bucket_size}
    tmp = "%s.%d.tmp.npz" % (side, os.getpid())
    try:
        np.savez(tmp, meta=np.frombuffer(json.dumps(meta).encode("utf-8"), dtype=np.uint8), chrom_id=idx.chrom_id,
                 L=idx.L, R=idx.R, ensg_id=idx.ensg_id, type_code=idx.type_code, strand_code=idx.strand_code)
        os.replace(tmp, side)
    except OSError:                                     # read-only directory: the cache is optional
        try:
            os.unlink(tmp)
        except OSError:
            pass


def load_glb(path, verify=True, cache=None):
    """Equivalent of miniglbase.glload (utils.py:21-59) + the flattening above.

    SURVEY.md 8f-3: unpickling a genome-scale index costs ~8.5 us per feature (about 50 s for hg38).
    With `cache` (default: on unless TEC_INDEX_CACHE=0) the flattened arrays are kept in a sidecar
    `<path>.tecidx.npz` next to the index.  The sidecar is derived data only: it is written after a
    successful, bucket-verified load of the .glb and is used only while the size and BLAKE2b digest
    of the .glb recorded in it match the file on disk; otherwise the .glb is read again."""
    real = os.path.realpath(path)
    assert os.path.exists(real), "File '%s' not found" % path
    use_cache = _cache_enabled(cache) and verify
    if use_cache:
        side = real + CACHE_SUFFIX
        fp = _fingerprint(real)
        if os.path.exists(side):
            idx = _load_sidecar(side, fp)
            if idx is not None:
                return idx
    idx = _load_glb_pickle(path, verify)
    if use_cache:
        _write_sidecar(side, fp, idx)
    return idx


def _load_glb_pickle(path, verify=True):
    gl = _read_pickle(path)
    rows = getattr(gl, "linearData", None)
    if rows is None:
        raise ValueError("%s is not a glbase genelist pickle (no linearData)" % path)
    buckets = getattr(gl, "buckets", None)
    if buckets is None:
        # utils.py:52-54: old lists are re-bucketed at load with the same closed form
        verify = False
    return from_rows(rows, buckets=buckets, verify=verify)
