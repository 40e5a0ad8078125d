"""Opt-in extensions: defined semantics where the reference raises.  CPU side: the rules of
oracle/te_oracle_ext.py on hand-checked cases, and the measureTE mirror with the oracle plugged in as the engine --
off by default (the reference's exceptions), on with measureTE(extensions=True)."""
import numpy as np
import pytest

import helpers as H
from oracle import te_oracle, te_oracle_ext
from oracle.ref_runner import CaptureLog
from oracle_engine import OracleEngine
import te_counter_b200
from te_counter_b200 import index as tindex
from test_host_mirror import install_stub_pysam

# the toy index of SURVEY.md Appendix A
ROWS = [
    {"loc": {"chr": "chr1", "left": 1000, "right": 2000}, "strand": "+", "type": "protein_coding", "ensg": "ENSG1"},
    {"loc": {"chr": "chr1", "left": 1500, "right": 1800}, "strand": "-", "type": "TE", "ensg": "LINE:L1:L1Md"},
    {"loc": {"chr": "chr1", "left": 5000, "right": 5300}, "strand": "-", "type": "TE", "ensg": "SINE:Alu:B1"},
    {"loc": {"chr": "chr1", "left": 20000, "right": 20500}, "strand": "+", "type": "lncRNA", "ensg": "ENSG2"},
    {"loc": {"chr": "chrX", "left": 9990, "right": 10010}, "strand": "+", "type": "pseudogene", "ensg": "ENSG3"},
]


def toy():
    idx = tindex.from_rows(ROWS)
    return idx, H.oracle_index(idx)


def test_defined_quality():
    assert te_oracle_ext.defined_quality([30]) == 30 and te_oracle_ext.defined_quality(20) == 20
    mte = te_counter_b200.measureTE("x", [30], extensions=True)
    assert mte._qual() == 30
    with pytest.raises(TypeError):                                   # bin/te_count:30 + te_count.py:88, as the reference
        te_counter_b200.measureTE("x", [30])._qual()
    with pytest.raises(TypeError):
        te_counter_b200.measureTE("x", [30, 40], extensions=True)._qual()


def test_bulk_stranded_rule_on_the_toy_index():
    idx, oidx = toy()
    names = idx.names
    c1 = idx.chrom_lookup["chr1"]
    # forward read over ENSG1 (+) and L1Md (-); the same read reversed; a reversed read over Alu (-); a forward one
    start = [1600, 1600, 5100, 5100]
    end = [1650, 1650, 5150, 5150]
    flag = [0, te_oracle.F_REVERSE, te_oracle.F_REVERSE, 0]
    chrom = [c1] * 4
    mapq = [60] * 4
    plain, _ = te_oracle.bulk_count(oidx, False, 20, start, end, chrom, mapq, flag)
    got, st = te_oracle_ext.bulk_count_stranded(oidx, False, 20, start, end, chrom, mapq, flag)
    assert dict(zip(names, plain)) == {"ENSG1": 2, "ENSG2": 0, "ENSG3": 0, "LINE:L1:L1Md": 2, "SINE:Alu:B1": 2}
    assert dict(zip(names, got)) == {"ENSG1": 1, "ENSG2": 0, "ENSG3": 0, "LINE:L1:L1Md": 1, "SINE:Alu:B1": 1}
    assert st["assigned"] == 3 and st["total_reads"] == 5            # the forward read over Alu hits nothing
    # paired end: the strand of the pair is that of its first record
    got_pe, _ = te_oracle_ext.bulk_count_stranded(oidx, True, 20, [1600, 1700], [0, 0], [c1, c1], [60, 60], [te_oracle.F_REVERSE, 0])
    assert dict(zip(names, got_pe))["LINE:L1:L1Md"] == 1 and dict(zip(names, got_pe))["ENSG1"] == 0


def _mirror(monkeypatch, records, qual=20, extensions=True):
    case = H.load_case(H.case_names("bulk_se")[0])
    mte = te_counter_b200.measureTE("test", qual, extensions=extensions)
    mte.bind_genome(H.GOLD + "/" + case["glb"])
    monkeypatch.setattr(mte, "_engine_obj", OracleEngine(0))
    install_stub_pysam(monkeypatch, H.stub_reads(records))
    mte.load_genome()
    return mte, case


def test_mirror_bulk_strand_extension(monkeypatch):
    case = H.load_case(H.case_names("bulk_se")[0])
    recs = [dict(r, flag=(r.get("flag", 0) | (0x10 if i % 3 == 0 else 0))) for i, r in enumerate(case["records"])]
    mte, _ = _mirror(monkeypatch, recs, case["qual"])
    res = mte.parse_bamse("mem.bam", strand=True, log=CaptureLog())
    idx = H.load_index(case["glb"])
    arr = H.pack_bulk(dict(case, records=recs), idx)
    want, st = te_oracle_ext.bulk_count_stranded(H.oracle_index(idx), False, case["qual"], *[arr[k].tolist() for k in ("start", "end", "chrom", "mapq", "flag")])
    assert res == dict(zip(idx.names, want)) and mte.total_reads == st["total_reads"]
    plain, _ = te_oracle.bulk_count(H.oracle_index(idx), False, case["qual"], *[arr[k].tolist() for k in ("start", "end", "chrom", "mapq", "flag")])
    assert sum(want) < sum(plain)                                    # the rule bites on this case
    mte_off, _ = _mirror(monkeypatch, recs, case["qual"], extensions=False)
    with pytest.raises(NotImplementedError):                         # te_count.py:183-184
        mte_off.parse_bamse("mem.bam", strand=True, log=CaptureLog())


def test_mirror_noumi_extension(monkeypatch, tmp_path):
    name = H.case_names("sc")[0]
    case = H.load_case(name)
    recs = [{k: v for k, v in r.items() if k not in ("UB", "UR")} for r in case["records"]]      # no UMI tags at all
    wl = tmp_path / "wl.txt"
    wl.write_text("".join(b + "\n" for b in case["whitelist"]))
    mte = te_counter_b200.measureTE("test", case["qual"], extensions=True)
    mte.bind_genome(H.GOLD + "/" + case["glb"])
    monkeypatch.setattr(mte, "_engine_obj", OracleEngine(0))
    install_stub_pysam(monkeypatch, H.stub_reads(recs))
    res = mte.sc_parse_bamse("mem.bam", UMIS=False, whitelistfilename=str(wl), strand=case["strand"], log=CaptureLog(),
                             label="x", maxcells=case["maxcells"])
    idx = H.load_index(case["glb"])
    arr, wl_obj = H.pack_sc(case, idx)                               # the tagged records: same columns except the UMI
    out = te_oracle_ext.sc_count_noumi(H.oracle_index(idx), case["qual"], case["strand"], 10_000_000, case["maxcells"], 1000,
                                       *[arr[k].tolist() for k in ("start", "end", "chrom", "mapq", "flag", "cell")])
    id_to_bc = wl_obj.id_to_barcode
    want = {}
    for (e, c), v in out["triples"].items():
        want.setdefault(idx.names[e], {})[id_to_bc[c]] = v
    assert want and {k: v for k, v in res.items() if v} == want
    off = te_counter_b200.measureTE("test", case["qual"])
    with pytest.raises(ZeroDivisionError):                           # te_count.py:703
        off.sc_parse_bamse("mem.bam", UMIS=False, whitelistfilename=str(wl), log=CaptureLog(), label="x", maxcells=3)
