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
