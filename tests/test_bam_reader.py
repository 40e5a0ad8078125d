"""te_counter_b200/bam.py (the stand-in for pysam) against files produced by an independent writer,
and the golden cases counted end to end from real BAM / SAM files: file -> reader -> packing ->
engine (the oracle here, the CUDA library in test_gpu_*) -> TSV bytes of the unmodified reference."""
import sys

import pytest

import helpers as H
from bam_writer import write_bam, write_sam
from oracle.ref_runner import CaptureLog
from oracle_engine import OracleEngine
import te_counter_b200
from te_counter_b200 import bam


def _records(case):
    recs = []
    for i, r in enumerate(case["records"]):
        r = dict(r)
        r.setdefault("name", "r%d" % i)
        recs.append(r)
    return recs


@pytest.mark.parametrize("fmt", ["bam", "sam"])
def test_reader_roundtrip(tmp_path, fmt):
    case = H.load_case("sc_rand_det")
    recs = _records(case)[:4000]
    path = str(tmp_path / ("x." + fmt))
    (write_bam if fmt == "bam" else write_sam)(path, recs)
    f = bam.AlignmentFile(path, "r")
    got = list(f)
    f.close()
    assert len(got) == len(recs)
    for g, r in zip(got, recs):
        fl = r.get("flag", 0)
        assert (g.is_unmapped, g.is_reverse, g.is_qcfail, g.is_duplicate) == \
            (bool(fl & 4), bool(fl & 16), bool(fl & 512), bool(fl & 1024))
        assert g.mapping_quality == r.get("mapq", 60) and g.query_name == r["name"]
        assert g.reference_name == r["chrom"] and g.reference_start == r["start"]
        if fl & 4 or r["end"] <= r["start"]:
            assert g.reference_end is None
        else:
            assert g.reference_end == r["end"]
        tags = dict(g.get_tags())
        for t in ("CB", "CR", "UB", "UR"):
            assert tags.get(t) == r.get(t)
        assert tags["NH"] == 1


def _mte(monkeypatch, case, path):
    monkeypatch.setitem(sys.modules, "pysam", None)          # `import pysam` raises ImportError -> bam.py
    monkeypatch.setenv("TEC_BAM_DECODER", "python")          # (the native decoder has tests/test_fastbam.py)
    mte = te_counter_b200.measureTE("test", case["qual"])
    mte.bind_genome(H.GOLD + "/" + case["glb"])
    monkeypatch.setattr(mte, "_engine_obj", OracleEngine(0))
    return mte


@pytest.mark.parametrize("name", ["bulk_se_rand_a", "bulk_pe_rand_a", "bulk_pe_odd", "bulk_se_appendixA"])
@pytest.mark.parametrize("fmt", ["bam", "sam"])
def test_bulk_case_from_file(monkeypatch, tmp_path, name, fmt):
    case = H.load_case(name)
    if any(r["end"] <= r["start"] and not r.get("flag", 0) & 4 for r in case["records"]):
        pytest.skip("case has zero-length alignments a file cannot carry")
    path = str(tmp_path / ("x." + fmt))
    (write_bam if fmt == "bam" else write_sam)(path, _records(case))
    mte = _mte(monkeypatch, case, path)
    mte.load_genome()
    log = CaptureLog()
    res = (mte.parse_bampe if case["paired"] else mte.parse_bamse)(path, strand=False, log=log)
    assert res == case["expected"]["result"] and mte.total_reads == case["expected"]["total_reads"]
    out = tmp_path / "o.tsv"
    mte.save_result_bulk(res, str(out), log=log)
    assert out.read_text() == case["expected"]["tsv"]


@pytest.mark.parametrize("name", ["sc_appendixA", "sc_rand_det_strand", "sc_rand_amb_bundles"])
def test_sc_case_from_file(monkeypatch, tmp_path, name):
    case = H.load_case(name)
    path = str(tmp_path / "x.bam")
    write_bam(path, _records(case))
    mte = _mte(monkeypatch, case, path)
    wl = tmp_path / "wl.txt"
    wl.write_text("".join(w + "\n" for w in case["whitelist"]))
    log = CaptureLog()
    res = mte.sc_parse_bamse(path, UMIS=True, whitelistfilename=str(wl), strand=case["strand"], log=log, label="l",
                             maxcells=case["maxcells"], _bundle_keys=case["bundle_keys"], _pad=case["pad"])
    assert {k: v for k, v in dict(res).items() if v} == case["expected"]["result"]
    out = tmp_path / "o.tsv"
    mte.sc_save_result(res, str(out), maxcells=case["maxcells"], log=log)
    assert out.read_text() == case["expected"]["tsv"]
