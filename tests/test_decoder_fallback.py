"""measureTE's order of BAM decoders (device -> libtecbam -> pysam -> bam.py): when the device decoder
refuses a file, cannot open it or has no memory for its window, the call starts over with a host
reader and the result is the golden one; any other device error propagates."""
import sys

import pytest

import helpers as H
from bam_writer import write_bam
from oracle.ref_runner import CaptureLog
from oracle_engine import OracleEngine
import te_counter_b200
from te_counter_b200 import _lib


class _Dev:
    def __init__(self, eng, exc, push_first):
        self.eng, self.exc, self.push_first = eng, exc, push_first

    def bind(self, cm, wl=None):
        pass

    def count(self, mode, qual):
        if self.push_first:                         # a window was already counted when the refusal came
            import numpy as np
            z = [np.zeros(2, d) for d in (np.int32, np.int32, np.uint16, np.uint8, np.uint8)]
            if mode == 2:
                self.eng.sc_push(2, *z, np.zeros(2, np.uint32), np.zeros(2, np.uint64))
            else:
                self.eng.bulk_push(2, *z)
        raise self.exc

    def info(self):
        return {}

    def close(self):
        pass


class _EngineWithDeviceDecoder(OracleEngine):
    exc = None
    push_first = False
    opened = 0

    def bam_open(self, filename):
        type(self).opened += 1
        if isinstance(self.exc, OSError):
            raise self.exc
        return _Dev(self, self.exc, self.push_first)


def _mte(monkeypatch, case, exc, push_first=False):
    monkeypatch.setitem(sys.modules, "pysam", None)
    monkeypatch.setenv("TEC_BAM_DECODER", "auto")
    _EngineWithDeviceDecoder.exc, _EngineWithDeviceDecoder.push_first, _EngineWithDeviceDecoder.opened = exc, push_first, 0
    mte = te_counter_b200.measureTE("test", case["qual"])
    mte.bind_genome(H.GOLD + "/" + case["glb"])
    monkeypatch.setattr(mte, "_engine_obj", _EngineWithDeviceDecoder(0))
    return mte


@pytest.mark.parametrize("exc,push_first", [(_lib.BamUnsupported("x"), False), (_lib.BamUnsupported("x"), True),
                                            (OSError("cannot open"), False),
                                            (_lib.TecError(_lib.ERR_NOMEM, "cudaMalloc: out of memory"), True)])
def test_bulk_starts_over_on_the_host(monkeypatch, tmp_path, exc, push_first):
    case = H.load_case("bulk_pe_rand_a")
    path = str(tmp_path / "x.bam")
    write_bam(path, [dict(r, name=r.get("name", "r%d" % i)) for i, r in enumerate(case["records"])])
    mte = _mte(monkeypatch, case, exc, push_first)
    mte.load_genome()
    res = mte.parse_bampe(path, strand=False, log=CaptureLog())
    assert _EngineWithDeviceDecoder.opened == 1
    assert res == case["expected"]["result"] and mte.total_reads == case["expected"]["total_reads"]


def test_sc_starts_over_on_the_host(monkeypatch, tmp_path):
    case = H.load_case("sc_rand_det")
    path = str(tmp_path / "x.bam")
    write_bam(path, [dict(r, name=r.get("name", "r%d" % i)) for i, r in enumerate(case["records"])])
    wl = tmp_path / "wl.txt"
    wl.write_text("".join(w + "\n" for w in case["whitelist"]))
    mte = _mte(monkeypatch, case, _lib.BamUnsupported("x"), True)
    res = mte.sc_parse_bamse(path, UMIS=True, whitelistfilename=str(wl), strand=case["strand"], log=CaptureLog(), label="l",
                             maxcells=case["maxcells"], _bundle_keys=case["bundle_keys"], _pad=case["pad"])
    assert {k: v for k, v in dict(res).items() if v} == case["expected"]["result"]


def test_other_device_errors_propagate(monkeypatch, tmp_path):
    case = H.load_case("bulk_pe_rand_a")
    path = str(tmp_path / "x.bam")
    write_bam(path, [dict(r, name=r.get("name", "r%d" % i)) for i, r in enumerate(case["records"])])
    # only the status TEC_ERR_NOMEM means "no room": other CUDA errors propagate, whatever their text says
    for exc in (_lib.TecError(-1, "an illegal memory access was encountered"), _lib.TecError(-1, "x: out of memory")):
        mte = _mte(monkeypatch, case, exc)
        mte.load_genome()
        with pytest.raises(_lib.TecError):
            mte.parse_bampe(path, strand=False, log=CaptureLog())
    mte = _mte(monkeypatch, case, AssertionError("CB or CR tag not found!"))
    mte.load_genome()
    with pytest.raises(AssertionError):
        mte.parse_bampe(path, strand=False, log=CaptureLog())
