"""Indices and inputs that leave the cell-table fast paths: more ensg than a 16-bit slot holds, an
ensg that carries two feature types, a bucket size other than 10000, empty indices / chromosomes,
degenerate features and negative read positions.  Every case must still match the oracle bit for
bit (the exact-search kernels take over where the table cannot be built)."""
import numpy as np
import pytest

import helpers as H
from oracle import te_oracle
from te_counter_b200 import _lib, synth
from te_counter_b200.index import GlbIndex

pytestmark = pytest.mark.gpu

SC_COLS = ("start", "end", "chrom", "mapq", "flag", "cell", "umi")


@pytest.fixture(scope="module")
def engine():
    eng = _lib.Engine(0)
    yield eng
    eng.close()


def bulk_check(engine, idx, r, paired, qual=20):
    engine.upload_index(idx)
    engine.bulk_begin(paired, qual)
    engine.bulk_push(len(r["start"]), r["start"], r["end"], r["chrom"], r["mapq"], r["flag"])
    counts, st = engine.bulk_finish()
    oc, os_ = te_oracle.bulk_count(H.oracle_index(idx), paired, qual, *[r[k].tolist() for k in ("start", "end", "chrom", "mapq", "flag")])
    assert counts.tolist() == oc
    assert (int(st[_lib.BS_UNITS]) + 1, st[_lib.BS_ASSIGNED], st[_lib.BS_LOWQ], st[_lib.BS_BADCHROM], st[_lib.BS_QCFAIL]) == \
        (os_["total_reads"], os_["assigned"], os_["lowq"], os_["badchrom"], os_["qcfail"])
    return counts


def sc_check(engine, idx, r, strand, n_wl, bundle_keys=10_000_000, maxcells=30, pad=10):
    engine.upload_index(idx)
    engine.sc_begin(20, strand, n_wl)
    engine.sc_push(len(r["start"]), *[r[k] for k in SC_COLS])
    nt, nh = engine.sc_finalize(bundle_keys, maxcells, pad)
    ensg, cell, count, hcell, hcount, st = engine.sc_fetch(nt, nh)
    out = te_oracle.sc_count(H.oracle_index(idx), 20, strand, bundle_keys, maxcells, pad, *[r[k].tolist() for k in SC_COLS])
    assert {(int(e), int(c)): int(v) for e, c, v in zip(ensg, cell, count)} == out["triples"]
    assert list(zip(hcell.tolist(), hcount.tolist())) == sorted(out["cell_hits"])
    assert int(st[_lib.SS_VALID]) == out["stats"]["valid"] and int(st[_lib.SS_ASSIGNED]) == out["stats"]["assigned"]


def reindex(idx, **kw):
    d = dict(chrom_keys=idx.chrom_keys, chrom_id=idx.chrom_id, L=idx.L, R=idx.R, ensg_id=idx.ensg_id,
             type_code=idx.type_code, strand_code=idx.strand_code, names=idx.names, bucket_size=idx.bucket_size)
    d.update(kw)
    out = GlbIndex(d["chrom_keys"], d["chrom_id"], d["L"], d["R"], d["ensg_id"], d["type_code"], d["strand_code"],
                   d["names"], bucket_size=d["bucket_size"])
    out.chrom_lengths = idx.chrom_lengths
    return out


def test_more_ensg_than_table_slots(engine):
    """70 000 distinct ensg: no 16-bit slots -> no cell table; the exact kernels count."""
    base = synth.synth_index(31, n_te=50000, n_exon=40000, n_gene=3000, chrom_len=6_000_000, n_chrom=2)
    rng = np.random.default_rng(1)
    n_names = 70000
    ensg = rng.integers(0, n_names, size=base.n_features).astype(np.int32)
    ensg[:n_names] = np.arange(n_names) if base.n_features >= n_names else ensg[:n_names]
    tcode = (1 + ensg % 2).astype(np.uint8)                       # type is a function of the ensg
    idx = reindex(base, ensg_id=ensg, type_code=tcode, names=["n%06d" % i for i in range(n_names)])
    r = synth.synth_bulk_reads(32, idx, 40000, paired=True, edge_frac=0.02)
    bulk_check(engine, idx, r, True)
    assert engine.get_info("has_stab") == 0 and engine.get_info("has_sc_stab") == 0
    rs = synth.synth_sc_reads(33, idx, 15000, n_whitelist=100, n_cells=30, umis_per_cell=30)
    sc_check(engine, idx, rs, True, 100)


def test_ensg_with_two_types(engine):
    """An ensg whose features have different types (e.g. a name used for a gene and a TE)."""
    base = synth.synth_index(34, n_te=20000, n_exon=6000, n_gene=400, chrom_len=2_000_000, n_chrom=2)
    tcode = base.type_code.copy()
    tcode[::7] = 0                                                # 'other' type on some rows of every ensg
    idx = reindex(base, type_code=tcode)
    r = synth.synth_bulk_reads(35, idx, 40000, paired=False, edge_frac=0.02)
    bulk_check(engine, idx, r, False)
    assert engine.get_info("has_stab") == 0
    rs = synth.synth_sc_reads(36, idx, 15000, n_whitelist=100, n_cells=30, umis_per_cell=30)
    sc_check(engine, idx, rs, False, 100)


@pytest.mark.parametrize("bs", [500, 4096, 12345])
def test_other_bucket_sizes(engine, bs):
    """miniglbase.config.bucket_size other than 10000: the generic edge test and candidate rule."""
    base = synth.synth_index(37, n_te=20000, n_exon=6000, n_gene=400, chrom_len=2_000_000, n_chrom=2)
    idx = reindex(base, bucket_size=bs)
    r = synth.synth_bulk_reads(38, idx, 40000, paired=True, edge_frac=0.0)
    # put many reads on the edges of THIS bucket size
    s = r["start"].copy()
    s[::5] = (s[::5] // bs) * bs
    s[1::5] = (s[1::5] // bs) * bs + bs - 1
    r["start"] = s
    r["end"] = (s + 100).astype(np.int32)
    bulk_check(engine, idx, r, True)
    bulk_check(engine, idx, r, False)
    rs = synth.synth_sc_reads(39, idx, 15000, n_whitelist=100, n_cells=30, umis_per_cell=30)
    sc_check(engine, idx, rs, True, 100)


def test_empty_index_and_empty_chromosome(engine):
    empty = GlbIndex(["1", "2"], np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32),
                     np.zeros(0, np.int32), np.zeros(0, np.uint8), np.zeros(0, np.uint8), [])
    r = {"start": np.array([5, 10, 20, 30], np.int32), "end": np.array([50, 60, 70, 80], np.int32),
         "chrom": np.array([0, 1, 0, 7], np.uint16), "mapq": np.full(4, 60, np.uint8), "flag": np.zeros(4, np.uint8)}
    bulk_check(engine, empty, r, False)
    bulk_check(engine, empty, r, True)
    # features on chromosome '3' only; reads on all three
    L = np.array([100, 5000, 20000], np.int32)
    idx = GlbIndex(["1", "2", "3"], np.full(3, 2, np.int32), L, L + 400, np.array([0, 1, 0], np.int32),
                   np.array([2, 1, 2], np.uint8), np.zeros(3, np.uint8), ["a", "b"])
    r = {"start": np.array([150, 150, 150, 5100, 20100, 0], np.int32), "end": np.array([250, 250, 250, 5200, 20200, 101], np.int32),
         "chrom": np.array([0, 1, 2, 2, 2, 2], np.uint16), "mapq": np.full(6, 60, np.uint8), "flag": np.zeros(6, np.uint8)}
    c = bulk_check(engine, idx, r, False)
    assert c.tolist() == [3, 1]


def test_degenerate_features_and_negative_positions(engine):
    """Zero-length and inverted features (R <= L), features at position 0, reads with negative
    coordinates or end <= start."""
    L = np.array([0, 0, 10, 50, 50, 9999, 10000, 10000, 20000, 20000], np.int32)
    R = np.array([0, 30, 10, 40, 51, 10000, 10000, 9990, 20001, 19999], np.int32)
    n = len(L)
    idx = GlbIndex(["1"], np.zeros(n, np.int32), L, R, (np.arange(n) % 4).astype(np.int32),
                   np.array([2, 1, 2, 2, 1, 2, 2, 1, 2, 2], np.uint8)[(np.arange(n) % 4)], np.zeros(n, np.uint8),
                   ["w", "x", "y", "z"])
    starts = np.array([-5, -1, 0, 0, 1, 9, 10, 11, 39, 40, 49, 50, 51, 9998, 9999, 10000, 10001, 19998, 19999, 20000, 20001, 29, 30], np.int32)
    ends = np.array([-1, 0, 0, 1, 2, 10, 11, 10, 41, 41, 50, 51, 50, 9999, 10000, 10001, 10000, 19999, 20000, 20001, 20002, 30, 31], np.int32)
    m = len(starts)
    r = {"start": starts, "end": ends, "chrom": np.zeros(m, np.uint16), "mapq": np.full(m, 60, np.uint8), "flag": np.zeros(m, np.uint8)}
    bulk_check(engine, idx, r, False, qual=0)
    r2 = {k: np.concatenate([v, v[:1]]) if m % 2 else v for k, v in r.items()}
    bulk_check(engine, idx, r2, True, qual=0)
    rs = dict(r, cell=(np.arange(m) % 2).astype(np.uint32), umi=((np.arange(m, dtype=np.uint64) + np.uint64(1)) << np.uint64(20)))
    for algo in (0, 1):
        engine.set_option("sc_algo", algo)
        sc_check(engine, idx, rs, False, 2, maxcells=2, pad=0)
        sc_check(engine, idx, rs, True, 2, maxcells=2, pad=0)
    engine.set_option("sc_algo", -1)
