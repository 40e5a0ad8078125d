"""libtecbam (include/tecbam.h, te_counter_b200/csrc/bamdecode.cpp) against the Python packing it
replaces: the same BAM file decoded by bam.AlignmentFile + reads.fill_bulk / fill_sc and by
NativeBam.fill_bulk / fill_sc must give the same arrays, batch by batch, and the same exceptions
where the reference's read loop raises; then golden cases end to end through measureTE."""
import ctypes
import os
import re
import struct
import sys

import numpy as np
import pytest

import helpers as H
from bam_writer import write_bam, write_sam, _bgzf_block
from oracle.ref_runner import CaptureLog
from oracle_engine import OracleEngine
import te_counter_b200
from te_counter_b200 import bam, build, fastbam, reads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ("start", "end", "chrom", "mapq", "flag")


@pytest.fixture(scope="module", autouse=True)
def _built():
    build.build_bam()


def test_header_symbols_exported_and_bound():
    src = open(os.path.join(ROOT, "include", "tecbam.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(tbam_[a-z_0-9]+)\s*\(", src)))
    assert len(names) >= 12
    lib = ctypes.CDLL(fastbam.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libtecbam.so does not export %s" % n
    assert sorted(fastbam.SIGNATURES) == names
    assert fastbam.load().tbam_abi_version() == 1


def _python_batches(path, mode, cm, wl, qual, cap):
    f = bam.AlignmentFile(path, "r")
    out, more = [], True
    while more:
        b = reads.Batch(cap, sc=mode == "sc")
        if mode == "sc":
            more = reads.fill_sc(b, f, cm, wl, qual)
        else:
            more = reads.fill_bulk(b, f, cm, mode == "pe", qual)
        out.append((b, more))
    f.close()
    return out


def _native_batches(path, mode, cm, wl, qual, cap, threads):
    f = fastbam.NativeBam(path, threads=threads)
    f.bind(cm, wl)
    out, more = [], True
    while more:
        b = reads.Batch(cap, sc=mode == "sc")
        more = f.fill_sc(b, qual) if mode == "sc" else f.fill_bulk(b, mode == "pe", qual)
        out.append((b, more))
    n = f.counters()["records"]
    f.close()
    assert n == sum(b.n for b, _ in out)
    return out


def _canon_chrom(batches, n_index, bulk=False):
    """Chromosomes outside the index get fresh ids >= n_index in single-cell mode: in order of first use in
    the Python packing, in header order in the library.  Only their identity matters (reads.ChromMap), so
    compare after relabelling by first appearance.  bulk: "not in the index" is all that matters
    (te_count.py:100 / :216), every such id is the same."""
    seen = {}
    out = []
    for b, _ in batches:
        c = b.chrom[:b.n].astype(np.int64)
        lab = c.copy()
        if bulk:
            lab[c >= n_index] = reads.CHROM_INVALID
            out.append(lab)
            continue
        for i in np.flatnonzero((c >= n_index) & (c < reads.MAX_CHROM_IDS)).tolist():
            lab[i] = seen.setdefault(int(c[i]), 1000000 + len(seen))
        out.append(lab)
    return out


def _same(py, nat, sc, n_index):
    assert [(b.n, m) for b, m in py] == [(b.n, m) for b, m in nat]
    for x, y in zip(_canon_chrom(py, n_index), _canon_chrom(nat, n_index)):
        assert np.array_equal(x, y), "chrom"
    for (a, _), (b, _) in zip(py, nat):
        for c in COLS + (("cell", "umi") if sc else ()):
            if c != "chrom":
                assert np.array_equal(getattr(a, c)[:a.n], getattr(b, c)[:b.n]), c


def _mixed_records(n, seed, sc):
    """Every kind of record the packing distinguishes: no reference, unknown contigs, skipped
    contigs, all four flag bits, low MAPQ, no CIGAR, mate names with and without '/', CB/CR and
    UB/UR fallbacks, barcodes outside the whitelist, UMIs with N."""
    rng = np.random.default_rng(seed)
    chroms = ["chr1", "chr2", "2", "chrX", "chrUn_KI270", "chr9_alt", "scaffold7", None]
    wl = ["ACGT%04d" % i for i in range(50)] + ["", "TTTTGGGG-1"]
    recs = []
    for i in range(n):
        ch = chroms[int(rng.integers(len(chroms)))]
        start = int(rng.integers(0, 150000))
        span = int(rng.choice([0, 1, 29, 30, 100, 5000]))
        flag = 0
        for bit, p in ((0x4, 0.05), (0x10, 0.5), (0x200, 0.05), (0x400, 0.05), (0x100, 0.1), (0x1, 0.5)):
            if rng.random() < p:
                flag |= bit
        if ch is None:
            flag |= 0x4
        if not sc and span == 0 and not flag & 0x4:
            span = 50                                   # an SE record without CIGAR raises (tested apart)
        if sc and span == 0 and ch not in ("chrUn_KI270", "chr9_alt"):
            span = 50
        r = {"chrom": ch, "start": start if ch is not None else -1, "end": start + span, "flag": flag,
             "mapq": int(rng.choice([0, 3, 19, 20, 60, 255]))}
        k = i // 2
        r["name"] = ["q%d" % k, "q%d/%d" % (k, i % 2 + 1), "lane/q%d/%d" % (k + (i % 2) * (k % 3 == 0), i % 2 + 1),
                     "lane_x/%d" % (i % 2) if i % 2 else "lane/x/%d" % (i % 2)][int(rng.integers(4))]
        if sc:
            bc = wl[int(rng.integers(len(wl)))] if rng.random() < 0.8 else "NOTINLIST"
            r["CB" if rng.random() < 0.6 else "CR"] = bc
            if rng.random() < 0.2:
                r["CR"] = "ACGT0001"                    # CB wins when both are there
            umi = "".join("ACGNT"[int(x)] for x in rng.integers(0, 5, int(rng.integers(0, 13))))
            r["UB" if rng.random() < 0.6 else "UR"] = umi
            if rng.random() < 0.2:
                r["UR"] = "AAAA"
        recs.append(r)
    return recs, wl


@pytest.mark.parametrize("mode", ["se", "pe", "sc"])
@pytest.mark.parametrize("block,cap,threads", [(3000, 1 << 16, 1), (700, 64, 3), (65000, 1001, 4), (64, 7, 2)])
def test_native_equals_python_packing(tmp_path, mode, block, cap, threads):
    recs, wl_list = _mixed_records(5000 + (mode == "pe"), 11 + block, mode == "sc")
    path = str(tmp_path / "x.bam")
    write_bam(path, recs, block=block)
    idx = H.load_index("idx_rand_a.glb")
    wl = None
    if mode == "sc":
        wlf = tmp_path / "wl.txt"
        wlf.write_text("".join(w + "\n" for w in wl_list))
        wl = reads.Whitelist(str(wlf))
    py = _python_batches(path, mode, reads.ChromMap(idx.chrom_keys), wl, 20, cap)
    nat = _native_batches(path, mode, reads.ChromMap(idx.chrom_keys), wl, 20, cap, threads)
    _same(py, nat, mode == "sc", len(idx.chrom_keys))
    assert sum(b.n for b, _ in nat) == (5000 if mode != "se" else 5000)


@pytest.mark.parametrize("window", [None, 1, 5000, 200000])
def test_large_file_many_windows(monkeypatch, tmp_path, window):
    """More records than one parse task and blocks of full size; records straddle blocks and, with
    a small decode window, windows (the cut record is carried over)."""
    if window:
        monkeypatch.setenv("TEC_BAM_WINDOW", str(window))
    case = H.load_case("sc_rand_det")
    recs = [dict(r, name="r%d" % i) for i, r in enumerate(case["records"])] * 3
    path = str(tmp_path / "x.bam")
    write_bam(path, recs, block=65280)
    idx = H.load_index(case["glb"])
    wlf = tmp_path / "wl.txt"
    wlf.write_text("".join(w + "\n" for w in case["whitelist"]))
    wl = reads.Whitelist(str(wlf))
    py = _python_batches(path, "sc", reads.ChromMap(idx.chrom_keys), wl, case["qual"], 1 << 15)
    nat = _native_batches(path, "sc", reads.ChromMap(idx.chrom_keys), wl, case["qual"], 1 << 15, 4)
    _same(py, nat, True, len(idx.chrom_keys))


def test_empty_file_and_references(tmp_path):
    path = str(tmp_path / "e.bam")
    write_bam(path, [])
    f = fastbam.NativeBam(path, threads=2)
    assert f.references == []
    f.bind(reads.ChromMap(["1"]))
    b = reads.Batch(16)
    assert f.fill_bulk(b, True, 20) is False and b.n == 0
    f.close()
    recs = [{"chrom": "chr7", "start": 5, "end": 50}, {"chrom": "chrM", "start": 7, "end": 60}]
    write_bam(path, recs)
    f = fastbam.NativeBam(path)
    assert f.references == bam.AlignmentFile(path).references == ["chr7", "chrM"]
    f.close()


def test_not_bgzf(tmp_path):
    p = str(tmp_path / "x.sam")
    write_sam(p, [{"chrom": "chr1", "start": 5, "end": 50}])
    with pytest.raises(fastbam.NotBgzf):
        fastbam.NativeBam(p)
    import gzip
    g = str(tmp_path / "plain.bam")
    with gzip.open(g, "wb") as fh:
        fh.write(b"BAM\1" + struct.pack("<ii", 0, 0))
    with pytest.raises(fastbam.NotBgzf):
        fastbam.NativeBam(g)
    with pytest.raises(OSError):
        fastbam.NativeBam(str(tmp_path / "missing.bam"))


def _one(tmp_path, recs, mode, wl_list=("AAAA",), qual=20):
    path = str(tmp_path / "x.bam")
    write_bam(path, recs)
    idx = H.load_index("idx_toy.glb")
    wl = None
    if mode == "sc":
        wlf = tmp_path / "wl.txt"
        wlf.write_text("".join(w + "\n" for w in wl_list))
        wl = reads.Whitelist(str(wlf))
    out = []
    for fn in (_python_batches, lambda *a: _native_batches(*a, 2)):
        try:
            fn(path, mode, reads.ChromMap(idx.chrom_keys), wl, qual, 64)
            out.append(None)
        except Exception as e:          # noqa: BLE001 -- the class is what is compared
            out.append((type(e), str(e).split(" (record")[0]))
    return out


def test_same_exceptions_as_python_packing(tmp_path):
    ok = {"chrom": "chr1", "start": 1600, "end": 1650, "CB": "AAAA", "UB": "ACGT"}
    # no barcode tag on a record that passes the filters / on one that does not
    py, nat = _one(tmp_path, [ok, dict(ok, CB=None)], "sc")
    assert py == nat == (AssertionError, "CB or CR tag not found!")
    py, nat = _one(tmp_path, [ok, dict(ok, CB=None, mapq=3), dict(ok, CB=None, flag=0x400)], "sc")
    assert py is None and nat is None
    # no UMI tag: only looked at once the barcode is in the whitelist
    py, nat = _one(tmp_path, [ok, dict(ok, UB=None)], "sc")
    assert py == nat == (AssertionError, "UB or UR tag not found!")
    py, nat = _one(tmp_path, [ok, dict(ok, CB="CCCC", UB=None)], "sc")
    assert py is None and nat is None
    # UMI the code cannot hold
    for umi in ("ACGTX", "A" * 22, "acgt"):
        py, nat = _one(tmp_path, [dict(ok, UB=umi)], "sc")
        assert py[0] is nat[0] is ValueError
    # reference_end is None on a counted record
    py, nat = _one(tmp_path, [dict(ok, end=1600)], "sc")
    assert py[0] is nat[0] is TypeError
    py, nat = _one(tmp_path, [dict(ok, end=1600)], "se")
    assert py == nat and py[0] is TypeError
    py, nat = _one(tmp_path, [dict(ok, end=1600, mapq=3), dict(ok, end=1600, chrom="chrZZ")], "se")
    assert py is None and nat is None
    # ':' in a chromosome name (--sc only)
    py, nat = _one(tmp_path, [ok, dict(ok, chrom="HLA:A")], "sc")
    assert py[0] is nat[0] is ValueError
    py, nat = _one(tmp_path, [ok, dict(ok, chrom="HLA:A")], "se")
    assert py is None and nat is None


def test_corrupt_and_truncated(tmp_path):
    recs = [{"chrom": "chr1", "start": 100 + i, "end": 150 + i} for i in range(2000)]
    path = str(tmp_path / "x.bam")
    write_bam(path, recs, block=20000)
    raw = open(path, "rb").read()
    cm = reads.ChromMap(["1"])

    def run(data):
        p = str(tmp_path / "y.bam")
        open(p, "wb").write(data)
        f = fastbam.NativeBam(p, threads=2)
        f.bind(cm)
        b = reads.Batch(1 << 12)
        while f.fill_bulk(b, False, 0):
            pass
        f.close()

    run(raw)
    flipped = bytearray(raw)
    flipped[len(raw) // 2] ^= 0x55                      # inside some block's deflate stream
    with pytest.raises(ValueError):
        run(bytes(flipped))
    with pytest.raises((ValueError, EOFError)):
        run(raw[:len(raw) // 2])                        # cut inside a block
    # cut at a block border but inside a record
    blocks, o = [], 0
    while o < len(raw):
        n = struct.unpack_from("<H", raw, o + 16)[0] + 1
        blocks.append(raw[o:o + n])
        o += n
    with pytest.raises(EOFError):
        run(b"".join(blocks[:3]))
    # impossible block_size
    body = b"BAM\1" + struct.pack("<iii", 0, 1, 5) + b"chr1\0" + struct.pack("<i", 1000) + struct.pack("<i", 7) + b"x" * 7
    with pytest.raises(ValueError):
        run(_bgzf_block(body) + _bgzf_block(b""))


def _mte(monkeypatch, case, decoder):
    monkeypatch.setitem(sys.modules, "pysam", None)
    monkeypatch.setenv("TEC_BAM_DECODER", decoder)
    mte = te_counter_b200.measureTE("test", case["qual"])
    mte.bind_genome(H.GOLD + "/" + case["glb"])
    monkeypatch.setattr(mte, "_engine_obj", OracleEngine(0))
    return mte


@pytest.mark.parametrize("name", ["bulk_se_rand_a", "bulk_pe_rand_a", "bulk_pe_odd", "bulk_pe_appendixA"])
def test_bulk_case_through_native_decoder(monkeypatch, tmp_path, name):
    case = H.load_case(name)
    if any(r["end"] <= r["start"] and not r.get("flag", 0) & 4 for r in case["records"]):
        pytest.skip("case has zero-length alignments a file cannot carry")
    path = str(tmp_path / "x.bam")
    write_bam(path, [dict(r, name=r.get("name", "r%d" % i)) for i, r in enumerate(case["records"])])
    mte = _mte(monkeypatch, case, "native")
    mte.load_genome()
    log = CaptureLog()
    res = (mte.parse_bampe if case["paired"] else mte.parse_bamse)(path, strand=False, log=log)
    assert res == case["expected"]["result"] and mte.total_reads == case["expected"]["total_reads"]
    out = tmp_path / "o.tsv"
    mte.save_result_bulk(res, str(out), log=log)
    assert out.read_text() == case["expected"]["tsv"]


@pytest.mark.parametrize("name", ["sc_appendixA", "sc_rand_det_strand", "sc_rand_amb_bundles"])
def test_sc_case_through_native_decoder(monkeypatch, tmp_path, name):
    case = H.load_case(name)
    path = str(tmp_path / "x.bam")
    write_bam(path, [dict(r, name=r.get("name", "r%d" % i)) for i, r in enumerate(case["records"])])
    mte = _mte(monkeypatch, case, "native")
    wl = tmp_path / "wl.txt"
    wl.write_text("".join(w + "\n" for w in case["whitelist"]))
    log = CaptureLog()
    res = mte.sc_parse_bamse(path, UMIS=True, whitelistfilename=str(wl), strand=case["strand"], log=log, label="l",
                             maxcells=case["maxcells"], _bundle_keys=case["bundle_keys"], _pad=case["pad"])
    assert {k: v for k, v in dict(res).items() if v} == case["expected"]["result"]
    out = tmp_path / "o.tsv"
    mte.sc_save_result(res, str(out), maxcells=case["maxcells"], log=log)
    assert out.read_text() == case["expected"]["tsv"]


def test_decoder_selection(monkeypatch, tmp_path):
    from te_counter_b200 import te_count as tc
    monkeypatch.setitem(sys.modules, "pysam", None)
    p = str(tmp_path / "x.bam")
    write_bam(p, [{"chrom": "chr1", "start": 5, "end": 50}])
    s = str(tmp_path / "x.sam")
    write_sam(s, [{"chrom": "chr1", "start": 5, "end": 50}])
    monkeypatch.delenv("TEC_BAM_DECODER", raising=False)
    assert isinstance(tc._open_alignment(p), fastbam.NativeBam)
    assert isinstance(tc._open_alignment(s), bam.AlignmentFile)         # SAM text: not BGZF
    monkeypatch.setenv("TEC_BAM_DECODER", "python")
    assert isinstance(tc._open_alignment(p), bam.AlignmentFile)
    monkeypatch.setenv("TEC_BAM_DECODER", "pysam")
    with pytest.raises(ImportError):
        tc._open_alignment(p)
    monkeypatch.setenv("TEC_BAM_DECODER", "native")
    with pytest.raises(fastbam.NotBgzf):
        tc._open_alignment(s)


@pytest.mark.parametrize("mode", ["se", "pe", "sc"])
@pytest.mark.parametrize("threads", [1, 3])
def test_committed_bam_fixtures(mode, threads):
    """tests/golden/bam_mixed_*.bam against the committed arrays of the Python packing
    (tests/make_bam_golden.py)."""
    want = np.load(os.path.join(H.GOLD, "bam_mixed_expected.npz"))
    idx = H.load_index("idx_rand_a.glb")
    wl = reads.Whitelist(os.path.join(H.GOLD, "bam_mixed_whitelist.txt")) if mode == "sc" else None
    nat = _native_batches(os.path.join(H.GOLD, "bam_mixed_%s.bam" % mode), mode, reads.ChromMap(idx.chrom_keys), wl, 20, 1024, threads)
    b = nat[0][0]
    assert b.n == 600 and nat[0][1] is False

    class _B:
        pass
    w = _B()
    w.n = 600
    for k in COLS + (("cell", "umi") if mode == "sc" else ()):
        setattr(w, k, want["%s_%s" % (mode, k)])
    _same([(w, False)], nat, mode == "sc", len(idx.chrom_keys))
