"""The C restatement of the bulk loop (oracle/te_oracle_c.c, the full-size checker) against the
Python oracle and, through the golden cases, against the unmodified reference."""
import numpy as np
import pytest

import helpers as H
from oracle import te_oracle, te_oracle_c
from te_counter_b200 import synth
from te_counter_b200.index import GlbIndex


def c_index(idx):
    return te_oracle_c.Index(idx.chrom_id, idx.L, idx.R, idx.ensg_id, idx.type_code, idx.n_chrom, idx.n_ensg, idx.bucket_size)


@pytest.mark.parametrize("name", H.case_names("bulk"))
def test_c_oracle_golden(name):
    case = H.load_case(name)
    idx = H.load_index(case["glb"])
    r = H.pack_bulk(case, idx)
    counts, st = te_oracle_c.bulk_count(c_index(idx), case["paired"], case["qual"], r["start"], r["end"], r["chrom"],
                                        r["mapq"], r["flag"], threads=3)
    exp = case["expected"]
    assert dict(zip(idx.names, counts.tolist())) == exp["result"]
    s = H.bulk_expected_stats(exp)
    assert (int(st[0]) + 1, int(st[1]), int(st[2]), int(st[3]), int(st[4])) == \
        (s["total_reads"], s["assigned"], s["lowq"], s["badchrom"], s["qcfail"])


@pytest.mark.parametrize("paired", [False, True])
@pytest.mark.parametrize("bs", [10000, 777])
def test_c_oracle_matches_python_oracle(paired, bs):
    base = synth.synth_index(41, n_te=20000, n_exon=6000, n_gene=400, chrom_len=1_500_000, n_chrom=3)
    idx = GlbIndex(base.chrom_keys, base.chrom_id, base.L, base.R, base.ensg_id, base.type_code, base.strand_code,
                   base.names, bucket_size=bs)
    idx.chrom_lengths = base.chrom_lengths
    r = synth.synth_bulk_reads(42, idx, 30000, paired=paired, edge_frac=0.05)
    s = r["start"].copy()
    s[::7] = (s[::7] // bs) * bs                      # edges of this bucket size
    s[3::11] = -5                                     # negative positions
    r["start"] = s
    counts, st = te_oracle_c.bulk_count(c_index(idx), paired, 20, r["start"], r["end"], r["chrom"], r["mapq"], r["flag"], threads=4)
    oc, os_ = te_oracle.bulk_count(H.oracle_index(idx), paired, 20, *[r[k].tolist() for k in ("start", "end", "chrom", "mapq", "flag")])
    assert counts.tolist() == oc
    assert (int(st[0]) + 1, int(st[1]), int(st[2]), int(st[3]), int(st[4])) == \
        (os_["total_reads"], os_["assigned"], os_["lowq"], os_["badchrom"], os_["qcfail"])


SC_COLS = ("start", "end", "chrom", "mapq", "flag", "cell", "umi")


def c_sc(idx, r, qual, strand, bundle_keys, maxcells, pad):
    feat = (idx.chrom_id, idx.L, idx.R, idx.ensg_id, idx.type_code, idx.strand_code)
    return te_oracle_c.sc_count(feat, idx.n_chrom, idx.bucket_size, qual, strand, bundle_keys, maxcells, pad,
                                *[r[k] for k in SC_COLS])


@pytest.mark.parametrize("name", [n for n in H.case_names("sc")])
def test_cpp_sc_oracle_golden(name):
    case = H.load_case(name)
    idx = H.load_index(case["glb"])
    r, wl = H.pack_sc(case, idx)
    out = c_sc(idx, r, case["qual"], case["strand"], case["bundle_keys"], case["maxcells"], case["pad"])
    exp = case["expected"]
    got = {}
    for (e, c), v in out["triples"].items():
        got.setdefault(idx.names[e], {})[wl.id_to_barcode[c]] = v
    assert got == exp["result"]
    assert [wl.id_to_barcode[c] for c, _ in out["cell_hits"]] == exp["barcode_order"]
    s = H.sc_expected_stats(exp)
    for k in ("total_reads", "invalid_barcode", "already_seen", "lowq", "qcfail", "valid", "assigned", "raw_barcodes"):
        assert out["stats"][k] == s[k], k


@pytest.mark.parametrize("strand", [False, True])
@pytest.mark.parametrize("bundle_keys,maxcells,pad", [(10_000_000, 50, 20), (300, 30, 10), (7, 200, 1000)])
def test_cpp_sc_oracle_matches_python_oracle(strand, bundle_keys, maxcells, pad):
    idx = synth.synth_index(43, n_te=20000, n_exon=6000, n_gene=400, chrom_len=1_500_000, n_chrom=3)
    r = synth.synth_sc_reads(44, idx, 12000, n_whitelist=150, n_cells=40, umis_per_cell=25)
    out = c_sc(idx, r, 20, strand, bundle_keys, maxcells, pad)
    ref = te_oracle.sc_count(H.oracle_index(idx), 20, strand, bundle_keys, maxcells, pad, *[r[k].tolist() for k in SC_COLS])
    assert out["triples"] == ref["triples"]
    assert out["cell_hits"] == ref["cell_hits"]
    for k in ("total_reads", "invalid_barcode", "already_seen", "lowq", "qcfail", "valid", "assigned", "raw_barcodes", "n_bundles"):
        assert out["stats"][k] == ref["stats"][k], k
