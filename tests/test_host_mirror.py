"""Host logic of the measureTE mirror (batching, string handling, log lines, TSV writers) on CPU,
with the oracle plugged in where the CUDA engine would be.  The GPU tests (test_gpu_*.py) run the
same cases through libtecount.so."""
import sys
import types

import pytest

import helpers as H
from oracle.ref_runner import CaptureLog
from oracle_engine import OracleEngine
import te_counter_b200
from te_counter_b200 import te_count as mirror


class _StubSam:
    def __init__(self, reads):
        self._it = iter(reads)

    def __iter__(self):
        return self

    def __next__(self):
        return next(self._it)

    def close(self):
        pass


def install_stub_pysam(monkeypatch, reads):
    m = types.ModuleType("pysam")
    m.AlignmentFile = lambda fn, mode="r": _StubSam(reads)
    monkeypatch.setitem(sys.modules, "pysam", m)


def make_mte(monkeypatch, case, engine_cls, batch=None):
    mte = te_counter_b200.measureTE("test", case["qual"])
    mte.bind_genome(H.GOLD + "/" + case["glb"])
    eng = engine_cls(0)
    monkeypatch.setattr(mte, "_engine_obj", eng)
    if batch:
        monkeypatch.setattr(mirror, "BATCH_RECORDS", batch)
    install_stub_pysam(monkeypatch, H.stub_reads(case["records"]))
    return mte


def run_bulk_case(monkeypatch, tmp_path, name, engine_cls, batch=None):
    case = H.load_case(name)
    mte = make_mte(monkeypatch, case, engine_cls, batch)
    mte.load_genome()
    log = CaptureLog()
    fn = mte.parse_bampe if case["paired"] else mte.parse_bamse
    res = fn("mem.bam", strand=False, log=log)
    exp = case["expected"]
    assert res == exp["result"]
    assert list(res.keys()) == sorted(exp["result"].keys())
    assert mte.total_reads == exp["total_reads"]
    out = tmp_path / "o.tsv"
    mte.save_result_bulk(res, str(out), log=log)
    assert out.read_text() == exp["tsv"]
    got = [m for _, m in log.lines if not m.startswith("Saved")]
    want = [m for m in exp["log"] if not m.startswith("Saved")]
    assert got == want


def run_sc_case(monkeypatch, tmp_path, name, engine_cls, batch=None):
    case = H.load_case(name)
    mte = make_mte(monkeypatch, case, engine_cls, batch)
    wl = tmp_path / "wl.txt"
    wl.write_text("".join(w + "\n" for w in case["whitelist"]))
    log = CaptureLog()
    res = mte.sc_parse_bamse("mem.bam", UMIS=True, whitelistfilename=str(wl), strand=case["strand"], log=log,
                             label="lbl", maxcells=case["maxcells"], _bundle_keys=case["bundle_keys"],
                             _pad=case["pad"])
    exp = case["expected"]
    assert {k: v for k, v in dict(res).items() if v} == exp["result"]
    assert list(res.keys()) == mte.all_feature_names
    assert mte.barcodes == exp["barcodes"] and list(mte.barcodes) == exp["barcode_order"]
    assert mte.total_reads == exp["total_reads"]
    out = tmp_path / "o.tsv"
    mte.sc_save_result(res, str(out), maxcells=case["maxcells"], log=log)
    assert out.read_text() == exp["tsv"]
    assert (tmp_path / "o.barcode_freq.tsv").read_text() == exp["freq"]
    # the generic (plain dict) writer path gives the same bytes
    out2 = tmp_path / "p.tsv"
    mte.sc_save_result({k: dict(v) for k, v in res.items()}, str(out2), maxcells=case["maxcells"], log=log)
    assert out2.read_text() == exp["tsv"]
    skip = ("Saved Bundle", "Consumed", "Cleaned up", "Waiting for all bundles", "Densifying", "Saving barcode",
            "Processed")
    got = [m for _, m in log.lines if not any(s in m for s in skip)]
    want = [m for m in exp["log"] if not any(s in m for s in skip)]
    assert got[:len(want)] == want


@pytest.mark.parametrize("name", H.case_names("bulk"))
def test_bulk_mirror_cpu(monkeypatch, tmp_path, name):
    run_bulk_case(monkeypatch, tmp_path, name, OracleEngine, batch=512)


@pytest.mark.parametrize("name", H.case_names("sc"))
def test_sc_mirror_cpu(monkeypatch, tmp_path, name):
    run_sc_case(monkeypatch, tmp_path, name, OracleEngine, batch=1000)


def test_crash_behaviours(monkeypatch, tmp_path):
    case = H.load_case("bulk_se_appendixA")
    mte = make_mte(monkeypatch, case, OracleEngine)
    mte.load_genome()
    with pytest.raises(NotImplementedError):
        mte.parse_bamse("mem.bam", strand=True, log=CaptureLog())
    mte.quality_threshold = [30]                       # bin/te_count:30 `-q 30` hands over a list
    with pytest.raises(TypeError):
        mte.parse_bamse("mem.bam", log=CaptureLog())
    with pytest.raises(AssertionError):
        mte.bind_genome("/nonexistent.glb")
    # PE names a/1, b/2 -> sys.quit AttributeError (te_count.py:92-94)
    pe = H.load_case("bulk_pe_appendixA")
    pe["records"][1]["name"] = "zz/2"
    mte = te_counter_b200.measureTE("test", 20)
    mte.bind_genome(H.GOLD + "/" + pe["glb"])
    monkeypatch.setattr(mte, "_engine_obj", OracleEngine(0))
    install_stub_pysam(monkeypatch, H.stub_reads(pe["records"]))
    mte.load_genome()
    with pytest.raises(AttributeError):
        mte.parse_bampe("mem.bam", log=CaptureLog())
