"""Parity of the CUDA single-cell path (through the C ABI) with the reference's outputs (golden
cases, incl. multi-bundle and hash-order-ambiguous inputs under the canonical rule) and with the
CPU oracle on larger seeded inputs."""
import numpy as np
import pytest

import helpers as H
from te_counter_b200 import synth
from oracle import te_oracle
from te_counter_b200 import _lib
from test_host_mirror import run_sc_case

pytestmark = pytest.mark.gpu

COLS = ("start", "end", "chrom", "mapq", "flag", "cell", "umi")


@pytest.fixture(scope="module")
def engine():
    eng = _lib.Engine(0)
    yield eng
    eng.close()


@pytest.mark.parametrize("name", H.case_names("sc"))
def test_sc_golden_through_mirror(monkeypatch, tmp_path, name):
    run_sc_case(monkeypatch, tmp_path, name, _lib.Engine, batch=1000)


def run_engine(engine, r, qual, strand, n_wl, bundle_keys, maxcells, pad, chunks=1):
    engine.sc_begin(qual, strand, n_wl)
    n = len(r["start"])
    step = (n + chunks - 1) // chunks if n else 1
    for a in range(0, max(n, 1), step):
        b = min(n, a + step)
        engine.sc_push(b - a, *[np.ascontiguousarray(r[k][a:b]) for k in COLS])
    nt, nh = engine.sc_finalize(bundle_keys, maxcells, pad)
    ensg, cell, count, hcell, hcount, st = engine.sc_fetch(nt, nh)
    sel = engine.sc_select(maxcells, nh)
    return ensg, cell, count, hcell, hcount, st, sel


def check_against_oracle(engine, idx, r, qual, strand, n_wl, bundle_keys, maxcells, pad, chunks=1):
    ensg, cell, count, hcell, hcount, st, sel = run_engine(engine, r, qual, strand, n_wl, bundle_keys, maxcells, pad, chunks)
    out = te_oracle.sc_count(H.oracle_index(idx), qual, strand, bundle_keys, maxcells, pad, *[r[k].tolist() for k in COLS])
    got = {(int(e), int(c)): int(v) for e, c, v in zip(ensg, cell, count)}
    assert got == out["triples"]
    assert list(zip(ensg.tolist(), cell.tolist())) == sorted(got)                 # sorted by (ensg, cell)
    assert list(zip(hcell.tolist(), hcount.tolist())) == sorted(out["cell_hits"])
    s = out["stats"]
    assert int(st[_lib.SS_UNITS]) + 1 == s["total_reads"]
    for k, f in ((_lib.SS_INVALID_BARCODE, "invalid_barcode"), (_lib.SS_ALREADY_SEEN, "already_seen"),
                 (_lib.SS_LOWQ, "lowq"), (_lib.SS_QCFAIL, "qcfail"), (_lib.SS_VALID, "valid"),
                 (_lib.SS_ASSIGNED, "assigned"), (_lib.SS_RAW_BARCODES, "raw_barcodes"), (_lib.SS_BUNDLES, "n_bundles")):
        assert int(st[k]) == s[f], f
    want_sel = sorted(out["cell_hits"], key=lambda t: (-t[1], t[0]))[:maxcells]
    assert sel.tolist() == [c for c, _ in want_sel]
    return out


@pytest.mark.parametrize("algo", [0, 1])
@pytest.mark.parametrize("strand", [False, True])
@pytest.mark.parametrize("bundle_keys,maxcells,pad", [(10_000_000, 50, 20), (700, 50, 20), (64, 20, 5), (5, 200, 1000)])
def test_sc_matches_oracle_seeded(engine, strand, bundle_keys, maxcells, pad, algo):
    """algo 0 = Part 3 by exact search only, 1 = cell table (exact search for degenerate fragments)."""
    idx = synth.synth_index(11, n_te=30000, n_exon=9000, n_gene=600, chrom_len=3_000_000, n_chrom=3)
    r = synth.synth_sc_reads(12, idx, 30000, n_whitelist=300, n_cells=60, umis_per_cell=40)
    engine.set_option("sc_algo", algo)
    engine.upload_index(idx)
    assert engine.get_info("has_sc_stab") == 1
    out = check_against_oracle(engine, idx, r, 20, strand, 300, bundle_keys, maxcells, pad, chunks=3)
    engine.set_option("sc_algo", -1)
    assert out["stats"]["assigned"] > 20


def test_sc_fragment_edge_geometry(engine):
    """Fragments that touch feature ends by exactly one base, sit on bucket and cell borders, start
    at 0, or are degenerate (end <= start): the 1-bp-wider inclusive tests of te_count.py:645/:648."""
    from te_counter_b200.index import GlbIndex
    L = np.array([0, 1000, 2047, 2048, 9999, 10000, 19990, 30000, 30000, 40960], np.int32)
    R = np.array([50, 2047, 2100, 4096, 10000, 10050, 20010, 30100, 30100, 45000], np.int32)
    n = len(L)
    idx = GlbIndex(["1"], np.zeros(n, np.int32), L, R, np.arange(n, dtype=np.int32) % 7,
                   np.array([1, 2, 2, 1, 2, 2, 1, 2, 2, 1], np.uint8), np.array([0, 1, 0, 1, 0, 1, 0, 1, 0, 1], np.uint8),
                   ["e%d" % i for i in range(7)])
    starts, ends = [], []
    for a, b in zip(L.tolist(), R.tolist()):
        for d in (-2, -1, 0, 1, 2):
            starts += [max(0, a + d), max(0, b + d), max(0, a + d - 90), max(0, b + d)]
            ends += [max(0, a + d) + 90, max(0, b + d) + 90, max(0, a + d), max(0, b + d)]      # the last one: end == start
    m = len(starts)
    r = {"start": np.array(starts, np.int32), "end": np.array(ends, np.int32), "chrom": np.zeros(m, np.uint16),
         "mapq": np.full(m, 60, np.uint8), "flag": np.zeros(m, np.uint8),
         "cell": (np.arange(m) % 3).astype(np.uint32), "umi": ((np.arange(m, dtype=np.uint64) + np.uint64(1)) << np.uint64(30))}
    for algo in (0, 1):
        engine.set_option("sc_algo", algo)
        engine.upload_index(idx)
        for strand in (False, True):
            check_against_oracle(engine, idx, r, 20, strand, 3, 10_000_000, 3, 0)
    engine.set_option("sc_algo", -1)


def test_sc_many_places_per_key(engine):
    """Keys seen on several chromosome:strand combinations with repeated and re-ordered fragments:
    exercises the first-inserted rule, the set semantics and 'later fragment wins' of Part 3."""
    rng = np.random.default_rng(5)
    idx = synth.synth_index(13, n_te=20000, n_exon=5000, n_gene=300, chrom_len=1_000_000, n_chrom=4)
    n = 20000
    feat = rng.integers(0, idx.n_features, size=n)
    place = rng.integers(0, 40, size=n)                     # few places -> many repeats
    pf = rng.integers(0, idx.n_features, size=40)
    start = (idx.L[pf[place]] + rng.integers(-3, 4, size=n)).clip(0).astype(np.int32)
    r = {"start": start, "end": (start + 90).astype(np.int32), "chrom": idx.chrom_id[pf[place]].astype(np.uint16),
         "mapq": np.full(n, 60, np.uint8), "flag": (rng.integers(0, 2, size=n) * 8).astype(np.uint8),
         "cell": rng.integers(0, 12, size=n).astype(np.uint32),
         "umi": (rng.integers(1, 30, size=n).astype(np.uint64) << np.uint64(40))}
    del feat
    engine.upload_index(idx)
    for strand in (False, True):
        for bk in (10_000_000, 50):
            check_against_oracle(engine, idx, r, 20, strand, 12, bk, 8, 2)


def test_sc_empty_and_all_filtered(engine):
    idx = synth.synth_index(11, n_te=3000, n_exon=900, n_gene=60, chrom_len=300_000, n_chrom=2)
    engine.upload_index(idx)
    engine.sc_begin(20, False, 10)
    nt, nh = engine.sc_finalize(100, 5, 5)
    assert (nt, nh) == (0, 0)
    _, _, _, _, _, st = engine.sc_fetch(nt, nh)
    assert st[_lib.SS_UNITS] == 0 and st[_lib.SS_VALID] == 0
    r = synth.synth_sc_reads(3, idx, 500, n_whitelist=10, n_cells=5, umis_per_cell=10)
    r["mapq"][:] = 0
    ensg, cell, count, hcell, hcount, st, sel = run_engine(engine, r, 20, False, 10, 100, 5, 5)
    assert len(ensg) == 0 and len(hcell) == 0 and st[_lib.SS_UNITS] == 500 and st[_lib.SS_VALID] == 0
    assert st[_lib.SS_LOWQ] + st[_lib.SS_QCFAIL] == 500


def test_sc_push_chunking_invariance(engine):
    idx = synth.synth_index(11, n_te=30000, n_exon=9000, n_gene=600, chrom_len=3_000_000, n_chrom=3)
    r = synth.synth_sc_reads(14, idx, 50000, n_whitelist=500, n_cells=80, umis_per_cell=60)
    engine.upload_index(idx)
    a = run_engine(engine, r, 20, True, 500, 3000, 40, 10, chunks=1)
    b = run_engine(engine, r, 20, True, 500, 3000, 40, 10, chunks=7)
    for x, y in zip(a, b):
        assert (x == y).all()


def test_sc_umi_key_packing(engine):
    """Fixed-length ACGT UMIs sort on 2-bit/char 32-bit keys; UMIs with N or mixed lengths fall back
    to the 3-bit/char 64-bit keys.  Both must give the oracle's answer."""
    from te_counter_b200.reads import encode_umi
    rng = np.random.default_rng(7)
    idx = synth.synth_index(11, n_te=30000, n_exon=9000, n_gene=600, chrom_len=3_000_000, n_chrom=3)
    engine.upload_index(idx)
    r = synth.synth_sc_reads(15, idx, 20000, n_whitelist=200, n_cells=40, umis_per_cell=30)
    for pack in (1, 0):
        engine.set_option("sc_pack_umi", pack)
        check_against_oracle(engine, idx, r, 20, True, 200, 900, 30, 10)
    engine.set_option("sc_pack_umi", 1)
    # mixed alphabet / lengths: N inside, shorter UMIs, one-character UMIs
    pool = ["ACGTACGTAC", "ACGTNCGTAC", "ACG", "A", "T", "TTTTTTTTTTTT", "ACGTACGTACG", "NNNN", "CATG", "CATGA"]
    codes = np.array([encode_umi(u) for u in pool], dtype=np.uint64)
    r2 = dict(r)
    r2["umi"] = codes[rng.integers(0, len(pool), size=len(r["umi"]))]
    check_against_oracle(engine, idx, r2, 20, False, 200, 50, 30, 10)
    # order check of the packing itself: 3-bit codes and their string order agree with the 2-bit keys
    us = sorted("".join(rng.choice(list("ACGT"), size=8)) for _ in range(200))
    assert [encode_umi(u) for u in us] == sorted(encode_umi(u) for u in us)


def test_sc_sort_variants(engine):
    """The packed-key path in its three forms -- prev[] placed through a radix pass on the position (forced: the inputs
    of a test are far below the size where it switches on), prev[] stored in place, and the library sort in two
    stages -- over several bundles, with and without --strand, chunks of 1 and 3 tiles in csrc/radix.cuh."""
    idx = synth.synth_index(11, n_te=30000, n_exon=9000, n_gene=600, chrom_len=3_000_000, n_chrom=3)
    engine.upload_index(idx)
    r = synth.synth_sc_reads(16, idx, 60000, n_whitelist=300, n_cells=70, umis_per_cell=50)
    try:
        for prev_part, sort, chunk in ((2, 1, 1), (2, 1, 3), (0, 1, 1), (1, 0, 1)):
            engine.set_option("sc_prev_partition", prev_part)
            engine.set_option("sc_sort", sort)
            engine.set_option("sc_sort_chunk", chunk)
            check_against_oracle(engine, idx, r, 20, True, 300, 2500, 40, 10)
            check_against_oracle(engine, idx, r, 20, False, 300, 10_000_000, 40, 10)
    finally:
        engine.set_option("sc_prev_partition", 0)
        engine.set_option("sc_sort", 1)
        engine.set_option("sc_sort_chunk", 1)


def _text_from_triples(n_ensg, ensg, cell, count, cells, barcodes):
    """te_count.py:752-754 restated on the triples: '\\t'.join([barcode] + [str(count or 0) ...])."""
    m = {(int(e), int(c)): int(v) for e, c, v in zip(ensg, cell, count)}
    return "".join("\t".join([b] + [str(m.get((e, c), 0)) for e in range(n_ensg)]) + "\n"
                   for c, b in zip(cells, barcodes)).encode()


def _matrix_text(engine, cells, barcodes):
    import io
    n = engine.sc_matrix_text(cells, barcodes)
    out = io.BytesIO()
    engine.sc_matrix_write(out, n, chunk=4096 + 7)            # several chunks, odd size
    assert out.tell() == n
    return out.getvalue()


def test_sc_matrix_text_rows(engine):
    """tec_sc_matrix_text against the reference's row formatting: zeros between entries, counts of
    1 to 5 digits, rows without any entry, barcodes of different lengths, any row order/subset."""
    idx = synth.synth_index(3, n_te=3000, n_exon=800, n_gene=60, chrom_len=300_000, n_chrom=2)
    engine.upload_index(idx)
    n_wl = 40
    rng = np.random.default_rng(5)
    r = synth.synth_sc_reads(6, idx, 30000, n_whitelist=n_wl, n_cells=12, umis_per_cell=300)
    # one cell with 12,345 UMIs on one feature -> a 5-digit count; another with 150
    f = int(np.argmax(idx.R - idx.L))
    extra = {k: [] for k in COLS}
    for cellid, n_umi in ((3, 12346), (7, 151)):
        u = np.arange(n_umi, dtype=np.uint64)
        code = np.zeros(n_umi, dtype=np.uint64)
        for k in range(8):                                  # 8-character UMIs over A,C,G,T,N in base 5
            code = (code << np.uint64(3)) | ((u // np.uint64(5 ** (7 - k))) % np.uint64(5) + np.uint64(1))
        code <<= np.uint64(3 * (21 - 8))
        extra["start"].append(np.full(n_umi, idx.L[f] + 5, np.int32))
        extra["end"].append(np.full(n_umi, idx.L[f] + 6, np.int32))
        extra["chrom"].append(np.full(n_umi, idx.chrom_id[f], np.uint16))
        extra["mapq"].append(np.full(n_umi, 60, np.uint8))
        extra["flag"].append(np.zeros(n_umi, np.uint8))
        extra["cell"].append(np.full(n_umi, cellid, np.uint32))
        extra["umi"].append(code)
    r = {k: np.concatenate([np.asarray(r[k])] + extra[k]).astype(np.asarray(r[k]).dtype) for k in COLS}
    ensg, cell, count, hcell, hcount, st, sel = run_engine(engine, r, 20, False, n_wl, 10 ** 7, 30, 1000)
    assert count.max() >= 10000 and len(sel) >= 5
    names = ["BC%d" % i + "-1" * (i % 3) for i in range(n_wl)]
    names[int(sel[1])] = ""                                  # an empty barcode string is a legal whitelist line
    for cells in (sel.tolist(), sel.tolist()[::-1][:4], [int(sel[0])],
                  sel.tolist()[:3] + [c for c in range(n_wl) if c not in set(hcell.tolist())][:2]):   # rows with no entry
        got = _matrix_text(engine, cells, [names[c] for c in cells])
        assert got == _text_from_triples(idx.n_ensg, ensg, cell, count, cells, [names[c] for c in cells])
    assert engine.sc_matrix_text([], []) == 0
    with pytest.raises(Exception):
        engine.sc_matrix_text([0, 0], ["a", "b"])            # the same cell twice
    with pytest.raises(Exception):
        engine.sc_matrix_text([n_wl], ["a"])                 # outside the whitelist


def test_sc_matrix_text_without_triples(engine):
    idx = synth.synth_index(3, n_te=3000, n_exon=800, n_gene=60, chrom_len=300_000, n_chrom=2)
    engine.upload_index(idx)
    engine.sc_begin(20, False, 5)
    nt, nh = engine.sc_finalize(10 ** 7, 3, 1000)
    assert nt == 0
    got = _matrix_text(engine, [4, 1], ["x", "yy"])
    assert got == ("x" + "\t0" * idx.n_ensg + "\nyy" + "\t0" * idx.n_ensg + "\n").encode()


@pytest.mark.parametrize("name", ["sc_rand_det_strand", "sc_rand_amb_bundles"])
@pytest.mark.parametrize("decoder", ["native", "auto"])
def test_sc_from_bam_file(monkeypatch, tmp_path, name, decoder):
    """BAM file -> libtecbam -> pinned batches -> CUDA single-cell path -> matrix rows formatted on
    the device -> the reference's TSV bytes."""
    import sys
    import te_counter_b200
    from bam_writer import write_bam
    from oracle.ref_runner import CaptureLog
    case = H.load_case(name)
    path = str(tmp_path / "x.bam")
    write_bam(path, [dict(r, name=r.get("name", "r%d" % i)) for i, r in enumerate(case["records"])])
    wl = tmp_path / "wl.txt"
    wl.write_text("".join(w + "\n" for w in case["whitelist"]))
    monkeypatch.setitem(sys.modules, "pysam", None)
    monkeypatch.setenv("TEC_BAM_DECODER", decoder)
    mte = te_counter_b200.measureTE("test", case["qual"], device=0)
    mte.bind_genome(H.GOLD + "/" + case["glb"])
    log = CaptureLog()
    res = mte.sc_parse_bamse(path, UMIS=True, whitelistfilename=str(wl), strand=case["strand"], log=log, label="l",
                             maxcells=case["maxcells"], _bundle_keys=case["bundle_keys"], _pad=case["pad"])
    assert {k: v for k, v in dict(res).items() if v} == case["expected"]["result"]
    out = tmp_path / "o.tsv"
    mte.sc_save_result(res, str(out), maxcells=case["maxcells"], log=log)
    assert out.read_text() == case["expected"]["tsv"]
