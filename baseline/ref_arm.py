"""
The reference arm of bench.py: the UNMODIFIED reference (oaxiom/te_counter) timed on the host cores.

`install()` (called by __graft_entry__.build() in the build container, where /root/reference exists)
copies the reference's `te_count` package and `bin/` into baseline/_ref/ -- git-ignored, but it
travels to the GPU box with the snapshot like the built .so files.  The reference has no setup.py /
pyproject.toml, so `pip install --target` has nothing to build; a plain copy of the tree is the install.
Nothing under baseline/_ref is edited.

The reference imports `pysam` (te_count/te_count.py:11), which is not in this image: a stub module is
placed in sys.modules whose AlignmentFile yields in-memory read objects with exactly the attributes the
read loops touch (te_count.py:81-98, :204-214, :394-438).  The objects are built BEFORE the timed region
(BAM decoding is excluded on both arms).

The index: `measureTE.parse_bam*` reach into `genome.linearData` and `genome.buckets`
(te_count.py:68-73, :106-116).  For the 5.9 M-feature benchmark index the reference's own
`genelist.load_list()` needs minutes and ~6 GB, so `make_genelist()` fills a genuine
`miniglbase.genelist` instance with the same two structures `_optimiseData` builds
(miniglbase/genelist.py:367-380), vectorised with numpy; tests/test_ref_arm.py checks on a small index
that they are equal to what `load_list()` produces.  The counting loops run unmodified.
"""
import importlib
import os
import shutil
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("TE_REFERENCE_ROOT", "/root/reference")


def install():
    """Copy the reference tree into baseline/_ref (build container only).  Returns a one-line outcome."""
    src = os.path.join(REFERENCE_ROOT, "te_count")
    if not os.path.isfile(os.path.join(src, "te_count.py")):
        return "reference tree not present (GPU box: the prebuilt baseline/_ref is used)" if available() else "reference tree not present"
    os.makedirs(REF_DIR, exist_ok=True)
    for name in ("te_count", "bin"):
        dst = os.path.join(REF_DIR, name)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(REFERENCE_ROOT, name), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return "copied %s -> baseline/_ref (no setup.py / pyproject.toml in the reference: nothing for pip to build)" % REFERENCE_ROOT


def available():
    return os.path.isfile(os.path.join(REF_DIR, "te_count", "te_count.py"))


# ----------------------------------------------------------------------------- pysam stub
class Read:
    __slots__ = ("is_unmapped", "is_duplicate", "is_qcfail", "mapping_quality", "query_name",
                 "reference_name", "reference_start", "reference_end", "is_reverse", "_tags")

    def get_tags(self):
        return self._tags


_FILES = {}


class _AlignmentFile:
    def __init__(self, filename, mode="r"):
        self._it = iter(_FILES[filename])

    def __iter__(self):
        return self

    def __next__(self):
        return next(self._it)

    def close(self):
        pass


def load():
    """Import the reference package from baseline/_ref under the stub pysam; returns module te_count."""
    assert available(), "baseline/_ref is missing: run __graft_entry__.build() where /root/reference exists"
    if "pysam" not in sys.modules or not getattr(sys.modules["pysam"], "_te_stub", False):
        m = types.ModuleType("pysam")
        m.AlignmentFile = _AlignmentFile
        m._te_stub = True
        sys.modules["pysam"] = m
    for k in [k for k in sys.modules if k == "te_count" or k.startswith("te_count.")]:
        del sys.modules[k]
    sys.path.insert(0, REF_DIR)
    try:
        mod = importlib.import_module("te_count")
    finally:
        sys.path.remove(REF_DIR)
    import logging
    logging.getLogger("glbase3").setLevel(logging.ERROR)
    return mod


class NullLog:
    def info(self, m):
        pass

    warning = error = info


# ----------------------------------------------------------------------------- index
TYPE_NAMES = {0: "other", 1: "protein_coding", 2: "TE", 3: "snRNA", 4: "enhancer"}


def make_genelist(mod, idx):
    """A miniglbase.genelist with linearData (feature dicts with a `location`) and buckets as
    genelist._optimiseData lays them out (genelist.py:367-380), for the features of a GlbIndex."""
    mg = mod.miniglbase
    bs = mg.config.bucket_size
    gl = mg.genelist()
    names = idx.names
    keys = idx.chrom_keys
    location = mg.location
    strand = {0: "+", 1: "-"}
    rows = []
    for c, l, r, e, t, s in zip(idx.chrom_id.tolist(), idx.L.tolist(), idx.R.tolist(), idx.ensg_id.tolist(),
                                idx.type_code.tolist(), idx.strand_code.tolist()):
        rows.append({"loc": location(chr=keys[c], left=l, right=r), "strand": strand.get(s, "+"), "name": names[e],
                     "type": TYPE_NAMES.get(t, "other"), "ensg": names[e]})
    gl.linearData = rows
    # buckets[chr][b] = [feature indices in linearData order] for every b in range(L//bs*bs, (R+bs)//bs*bs, bs)
    L = idx.L.astype(np.int64)
    R = idx.R.astype(np.int64)
    lo = L // bs
    n_b = np.maximum((R + bs) // bs - lo, 0)
    feat = np.repeat(np.arange(len(L), dtype=np.int64), n_b)
    first = np.cumsum(n_b) - n_b
    b = (lo[feat] + (np.arange(len(feat), dtype=np.int64) - first[feat])) * bs
    chrom = idx.chrom_id.astype(np.int64)[feat]
    order = np.lexsort((feat, b, chrom))
    feat, b, chrom = feat[order], b[order], chrom[order]
    buckets = {}
    # every chromosome with a feature row is a key, even when all its bucket ranges are empty
    for c in np.unique(idx.chrom_id).tolist():
        buckets[keys[c]] = {}
    if len(feat):
        cut = np.flatnonzero((np.diff(chrom) != 0) | (np.diff(b) != 0)) + 1
        starts = np.concatenate([[0], cut])
        ends = np.concatenate([cut, [len(feat)]])
        fl = feat.tolist()
        for s0, e0 in zip(starts.tolist(), ends.tolist()):
            buckets[keys[int(chrom[s0])]][int(b[s0])] = fl[s0:e0]
    gl.buckets = buckets
    return gl


def new_measure(mod, gl, names, qual=20):
    mte = mod.measureTE("bench", qual)
    mte.genome = gl
    mte.all_feature_names = sorted(set(names))
    return mte


# ----------------------------------------------------------------------------- reads
def bulk_reads(idx, start, end, chrom, mapq, flag):
    """Stub read objects for SoA arrays (flag bits as in include/tecount.h)."""
    keys = ["chr" + k for k in idx.chrom_keys]
    out = []
    for i, (s, e, c, q, f) in enumerate(zip(start.tolist(), end.tolist(), chrom.tolist(), mapq.tolist(), flag.tolist())):
        r = Read()
        r.is_unmapped = bool(f & 1)
        r.is_duplicate = bool(f & 2)
        r.is_qcfail = bool(f & 4)
        r.is_reverse = bool(f & 8)
        r.mapping_quality = q
        r.query_name = "q"
        r.reference_name = keys[c] if c < len(keys) else "chrUn%d" % c
        r.reference_start = s
        r.reference_end = e
        r._tags = ()
        out.append(r)
    return out


_UMI_CH = " ACGNT"


def umi_string(code):
    s = []
    for i in range(21):
        d = (code >> (3 * (20 - i))) & 7
        if d == 0:
            break
        s.append(_UMI_CH[d])
    return "".join(s)


def sc_reads(idx, whitelist, start, end, chrom, mapq, flag, cell, umi):
    keys = ["chr" + k for k in idx.chrom_keys]
    out = []
    for s, e, c, q, f, b, u in zip(start.tolist(), end.tolist(), chrom.tolist(), mapq.tolist(), flag.tolist(),
                                   cell.tolist(), umi.tolist()):
        r = Read()
        r.is_unmapped = bool(f & 1)
        r.is_duplicate = bool(f & 2)
        r.is_qcfail = bool(f & 4)
        r.is_reverse = bool(f & 8)
        r.mapping_quality = q
        r.query_name = "q"
        r.reference_name = "chrUn_alt" if c == 0xFFFE else keys[c] if c < len(keys) else "chrZ%d" % c
        r.reference_start = s
        r.reference_end = e
        r._tags = (("CB", whitelist[b] if b < len(whitelist) else "NOTINLIST"), ("UB", umi_string(u)))
        out.append(r)
    return out


def run_bulk(mte, reads, paired):
    _FILES["mem.bam"] = reads
    return (mte.parse_bampe if paired else mte.parse_bamse)("mem.bam", strand=False, log=NullLog())


def run_sc(mte, reads, whitelist_path, strand, maxcells, workdir):
    """sc_parse_bamse writes its bundle files into the current directory (te_count.py:381) and reloads the
    .glb in Part 3 (te_count.py:580): run inside workdir, with load_genome bound to a no-op on the INSTANCE
    (the index object is already in place; no reference source is touched)."""
    _FILES["mem.bam"] = reads
    mte.load_genome = lambda: None
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        return mte.sc_parse_bamse("mem.bam", UMIS=True, whitelistfilename=whitelist_path, strand=strand, log=NullLog(),
                                  label="bench", maxcells=maxcells)
    finally:
        os.chdir(cwd)
