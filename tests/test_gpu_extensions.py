"""Opt-in extensions on the GPU (rules in oracle/te_oracle_ext.py, outside the parity claim):
strand-aware bulk counting in the exact-search kernel, and --noumi / -q N through measureTE on real BAM files."""
import numpy as np
import pytest

import helpers as H
from bam_writer import write_bam
from oracle import te_oracle, te_oracle_ext
from oracle.ref_runner import CaptureLog
import te_counter_b200
from te_counter_b200 import _lib, synth

pytestmark = pytest.mark.gpu
BULK = ("start", "end", "chrom", "mapq", "flag")


@pytest.fixture(scope="module")
def engine():
    eng = _lib.Engine(0)
    yield eng
    eng.close()


@pytest.mark.parametrize("paired", [False, True])
def test_bulk_strand_kernel_matches_its_restatement(engine, paired):
    idx = synth.synth_index(31, n_te=20000, n_exon=6000, n_gene=400, chrom_len=2_000_000, n_chrom=3)
    engine.upload_index(idx)
    r = synth.synth_bulk_reads(32, idx, 40000, paired=paired, edge_frac=0.05)
    rng = np.random.default_rng(5)
    flag = r["flag"].copy()
    flag[rng.random(len(flag)) < 0.5] |= _lib.F_REVERSE if hasattr(_lib, "F_REVERSE") else 8
    cols = [r["start"], r["end"], r["chrom"], r["mapq"], flag]
    oidx = H.oracle_index(idx)
    want, ws = te_oracle_ext.bulk_count_stranded(oidx, paired, 20, *[c.tolist() for c in cols])
    plain, _ = te_oracle.bulk_count(oidx, paired, 20, *[c.tolist() for c in cols])
    engine.set_option("bulk_strand", 1)
    try:
        engine.bulk_begin(paired, 20)
        engine.bulk_push(len(flag), *cols)
        counts, st = engine.bulk_finish()
    finally:
        engine.set_option("bulk_strand", 0)
    assert counts.tolist() == want and counts.tolist() != plain
    assert int(st[_lib.BS_ASSIGNED]) == ws["assigned"] and int(st[_lib.BS_UNITS]) + 1 == ws["total_reads"]
    engine.bulk_begin(paired, 20)                                    # and the option leaves nothing behind
    engine.bulk_push(len(flag), *cols)
    counts, _ = engine.bulk_finish()
    assert counts.tolist() == plain


def test_measure_te_extensions_on_files(tmp_path):
    case = H.load_case(H.case_names("bulk_se")[0])
    recs = [dict(r, flag=(r.get("flag", 0) | (0x10 if i % 3 == 0 else 0))) for i, r in enumerate(case["records"])]
    bam = tmp_path / "se.bam"
    write_bam(str(bam), recs)
    mte = te_counter_b200.measureTE("test", [case["qual"]], extensions=True)          # -q N as bin/te_count hands it over
    mte.bind_genome(H.GOLD + "/" + case["glb"])
    mte.load_genome()
    res = mte.parse_bamse(str(bam), strand=True, log=CaptureLog())
    idx = H.load_index(case["glb"])
    arr = H.pack_bulk(dict(case, records=recs), idx)
    want, st = te_oracle_ext.bulk_count_stranded(H.oracle_index(idx), False, case["qual"], *[arr[k].tolist() for k in BULK])
    assert res == dict(zip(idx.names, want)) and mte.total_reads == st["total_reads"]
    # --noumi on a file without UMI tags
    sc = H.load_case(H.case_names("sc")[0])
    recs = [{k: v for k, v in r.items() if k not in ("UB", "UR")} for r in sc["records"]]
    bam2 = tmp_path / "sc.bam"
    write_bam(str(bam2), recs)
    wl = tmp_path / "wl.txt"
    wl.write_text("".join(b + "\n" for b in sc["whitelist"]))
    mte = te_counter_b200.measureTE("test", sc["qual"], extensions=True)
    mte.bind_genome(H.GOLD + "/" + sc["glb"])
    res = mte.sc_parse_bamse(str(bam2), UMIS=False, whitelistfilename=str(wl), strand=sc["strand"], log=CaptureLog(),
                             label="x", maxcells=sc["maxcells"])
    idx = H.load_index(sc["glb"])
    arr, wl_obj = H.pack_sc(sc, idx)
    out = te_oracle_ext.sc_count_noumi(H.oracle_index(idx), sc["qual"], sc["strand"], 10_000_000, sc["maxcells"], 1000,
                                       *[arr[k].tolist() for k in BULK + ("cell",)])
    want = {}
    for (e, c), v in out["triples"].items():
        want.setdefault(idx.names[e], {})[wl_obj.id_to_barcode[c]] = v
    assert want and {k: dict(v) for k, v in res.items() if v} == want
