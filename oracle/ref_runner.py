"""
TEST INFRASTRUCTURE -- runs the UNMODIFIED reference (te_counter, /root/reference) in this
container, so that the restatement in oracle/te_oracle.py can be pinned against it and golden
vectors can be generated for tests/golden/.  Nothing here is on the product path.

/root/reference does not exist on the GPU box: this module is imported only by
oracle/make_golden.py and by tests that skip when the reference tree is absent.

The reference imports `pysam` at module top (te_count/te_count.py:11), which is not installed in
this image.  A ~30-line stub module is placed in sys.modules first; it yields in-memory read
objects carrying exactly the attributes the reference reads (te_count.py:81-98, 204-214, 394-438).

Two knobs that do not touch the reference's source tree on disk:
  * bundle_keys: the literal `1e7` at te_count.py:377 is replaced IN A TEMP COPY of the module
    source compiled under /tmp, to exercise multi-bundle logic at test scale.
  * pad: likewise the `+1000` of te_count.py:502 (cells kept for Part 2), so that the
    "non-selected barcode between two selected ones" branch of the Part-2 scan is reachable with a
    few dozen cells.
  * ordered_sets: the name `set` in the module globals of the loaded te_count.te_count is bound to
    an insertion-ordered set class.  The reference's `next(iter(umis[umi]))` (te_count.py:452)
    depends on PYTHONHASHSEED once a set has >=2 members; with insertion-ordered sets it becomes
    the canonical "first inserted fragment" rule the product and the oracle implement
    (SURVEY.md 8a-9).  No source line is changed for this.
"""
import importlib
import importlib.util
import os
import shutil
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("TE_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "te_count", "te_count.py"))


# ----------------------------------------------------------------------------- pysam stub
class StubRead:
    __slots__ = ("is_unmapped", "is_duplicate", "is_qcfail", "mapping_quality", "query_name",
                 "reference_name", "reference_start", "reference_end", "is_reverse", "_tags")

    def __init__(self, chrom, start, end, mapq=60, flag=0, name="r", tags=None):
        self.is_unmapped = bool(flag & 0x4)
        self.is_duplicate = bool(flag & 0x400)
        self.is_qcfail = bool(flag & 0x200)
        self.is_reverse = bool(flag & 0x10)
        self.mapping_quality = mapq
        self.query_name = name
        self.reference_name = chrom
        self.reference_start = start
        self.reference_end = end
        self._tags = tags or []

    def get_tags(self):
        return list(self._tags)


_REGISTRY = {}


class _StubAlignmentFile:
    def __init__(self, filename, mode="r"):
        self._it = iter(_REGISTRY[filename])

    def __iter__(self):
        return self

    def __next__(self):
        return next(self._it)

    def close(self):
        pass


def install_pysam_stub():
    if "pysam" not in sys.modules or not hasattr(sys.modules["pysam"], "_te_stub"):
        m = types.ModuleType("pysam")
        m.AlignmentFile = _StubAlignmentFile
        m._te_stub = True
        sys.modules["pysam"] = m


def register_reads(name, reads):
    _REGISTRY[name] = reads


# ----------------------------------------------------------------------------- ordered set
class OrderedSet:
    """Insertion-ordered set with the operations te_count.py uses on sets."""

    def __init__(self, it=()):
        self._d = dict.fromkeys(it)

    def add(self, x):
        self._d.setdefault(x, None)

    def update(self, it):
        for x in it:
            self._d.setdefault(x, None)

    def __contains__(self, x):
        return x in self._d

    def __iter__(self):
        return iter(self._d)

    def __len__(self):
        return len(self._d)

    def __bool__(self):
        return bool(self._d)

    def __or__(self, other):
        r = OrderedSet(self._d)
        r.update(other)
        return r


# ----------------------------------------------------------------------------- loading
_TMP_DIRS = []


def load_reference(bundle_keys=None, ordered_sets=False, pad=None):
    """Import the reference package and return it (module `te_count`).

    bundle_keys=None imports straight from REFERENCE_ROOT.  Otherwise the tree is copied to a
    temp dir and the single literal at te_count.py:377 is rewritten there.
    """
    assert reference_available(), "reference tree not present"
    install_pysam_stub()
    for k in [k for k in sys.modules if k == "te_count" or k.startswith("te_count.")]:
        del sys.modules[k]
    root = REFERENCE_ROOT
    if bundle_keys is not None or pad is not None:
        tmp = tempfile.mkdtemp(prefix="te_ref_")
        _TMP_DIRS.append(tmp)
        shutil.copytree(os.path.join(REFERENCE_ROOT, "te_count"), os.path.join(tmp, "te_count"))
        p = os.path.join(tmp, "te_count", "te_count.py")
        src = open(p).read()
        if bundle_keys is not None:
            needle = "if len(umis) >= 1e7:"
            assert src.count(needle) == 1
            src = src.replace(needle, "if len(umis) >= %d:" % int(bundle_keys))
        if pad is not None:                      # the +1000 of te_count.py:502
            needle = "[0:maxcells+1000]"
            assert src.count(needle) == 1
            src = src.replace(needle, "[0:maxcells+%d]" % int(pad))
        open(p, "w").write(src)
        root = tmp
    sys.path.insert(0, root)
    try:
        mod = importlib.import_module("te_count")
    finally:
        sys.path.remove(root)
    import logging
    logging.getLogger("glbase3").setLevel(logging.ERROR)
    if ordered_sets:
        mod.te_count.set = OrderedSet
    return mod


def cleanup():
    for d in _TMP_DIRS:
        shutil.rmtree(d, ignore_errors=True)
    _TMP_DIRS.clear()


class CaptureLog:
    """Stands in for the logging.Logger the CLI passes; keeps the messages."""

    def __init__(self):
        self.lines = []

    def info(self, m):
        self.lines.append(("info", m))

    def warning(self, m):
        self.lines.append(("warning", m))

    def error(self, m):
        self.lines.append(("error", m))


def build_index(mod, features, path):
    """features: list of dicts {chr,left,right,strand,type,ensg[,name]} -> .glb written by the
    reference's own genelist.load_list().save() (miniglbase/genelist.py:1429, base_genelist.py:302)."""
    mg = mod.miniglbase
    rows = []
    for f in features:
        rows.append({"loc": mg.location(chr=f["chr"], left=f["left"], right=f["right"]),
                     "strand": f["strand"], "name": f.get("name", f["ensg"]),
                     "type": f["type"], "ensg": f["ensg"]})
    gl = mg.genelist()
    gl.load_list(rows)
    gl.save(path)
    return path


def make_reads(recs, sc=False):
    """recs: list of dicts {chrom,start,end,mapq,flag[,name][,CB|CR][,UB|UR]}"""
    out = []
    for i, r in enumerate(recs):
        tags = []
        for t in ("CB", "CR", "UB", "UR"):
            if r.get(t) is not None:
                tags.append((t, r[t]))
        out.append(StubRead(r["chrom"], r["start"], r["end"], r.get("mapq", 60), r.get("flag", 0),
                            r.get("name", "r%d" % i), tags))
    return out


def run_bulk(mod, glb, recs, paired, qual=20):
    mte = mod.measureTE("oracle", qual)
    mte.bind_genome(glb)
    mte.load_genome()
    register_reads("mem.bam", make_reads(recs))
    log = CaptureLog()
    fn = mte.parse_bampe if paired else mte.parse_bamse
    res = fn("mem.bam", strand=False, log=log)
    with tempfile.NamedTemporaryFile("r", suffix=".tsv", delete=False) as t:
        tsv_path = t.name
    mte.save_result_bulk(res, tsv_path, log=log)
    tsv = open(tsv_path).read()
    os.remove(tsv_path)
    return {"result": res, "total_reads": mte.total_reads, "tsv": tsv,
            "log": [m for _, m in log.lines]}


def run_sc(mod, glb, recs, whitelist, maxcells, strand=False, qual=20, label="lbl"):
    mte = mod.measureTE("oracle", qual)
    mte.bind_genome(glb)
    register_reads("mem.bam", make_reads(recs, sc=True))
    log = CaptureLog()
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="te_sc_")
    os.chdir(tmp)          # the reference drops tmp.*.bun files in cwd (te_count.py:381)
    try:
        wl = os.path.join(tmp, "wl.txt")
        with open(wl, "w") as oh:
            for w in whitelist:
                oh.write(w + "\n")
        res = mte.sc_parse_bamse("mem.bam", UMIS=True, whitelistfilename=wl, strand=strand, log=log,
                                 label=label, maxcells=maxcells)
        barcodes_part3 = dict(mte.barcodes)
        out = os.path.join(tmp, "out.tsv")
        mte.sc_save_result(res, out, maxcells=maxcells, log=log)
        tsv = open(out).read()
        freq = open(out.replace(".tsv", ".barcode_freq.tsv")).read()
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)
    return {"result": {k: dict(v) for k, v in res.items() if v}, "barcodes": barcodes_part3,
            "barcode_order": list(barcodes_part3.keys()),
            "total_reads": mte.total_reads, "tsv": tsv, "freq": freq,
            "log": [m for _, m in log.lines]}
