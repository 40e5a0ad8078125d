"""Parity of the CUDA bulk path (through the C ABI) with the reference's outputs (golden cases)
and with the CPU oracle on larger seeded inputs."""
import numpy as np
import pytest

import helpers as H
from te_counter_b200 import synth
from oracle import te_oracle
from te_counter_b200 import _lib
from test_host_mirror import run_bulk_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    eng = _lib.Engine(0)
    yield eng
    eng.close()


@pytest.mark.parametrize("name", H.case_names("bulk"))
def test_bulk_golden_through_mirror(monkeypatch, tmp_path, name):
    run_bulk_case(monkeypatch, tmp_path, name, _lib.Engine, batch=1024)


@pytest.mark.parametrize("algo,shift", [(0, 11), (1, 11), (1, 9), (1, 8), (2, 10), (2, 9), (2, 8)])
@pytest.mark.parametrize("paired", [False, True])
@pytest.mark.parametrize("seed", [1, 2])
def test_bulk_matches_oracle_seeded(engine, paired, seed, algo, shift):
    """algo 0 = exact search kernel, 1 = cell-table kernel with in-kernel rings (round 1), 2 = two-pass kernels of
    bulk2.cuh (the default), several cell sizes."""
    idx = synth.synth_index(seed, n_te=30000, n_exon=9000, n_gene=600, chrom_len=3_000_000, n_chrom=3)
    r = synth.synth_bulk_reads(seed + 10, idx, 60000, paired=paired, edge_frac=0.1)
    engine.set_option("bulk_algo", algo)
    engine.set_option("stab_shift", shift)
    engine.upload_index(idx)
    assert engine.get_info("has_stab") == 1
    engine.bulk_begin(paired, 20)
    engine.bulk_push(len(r["start"]), r["start"], r["end"], r["chrom"], r["mapq"], r["flag"])
    counts, st = engine.bulk_finish()
    oc, os_ = te_oracle.bulk_count(H.oracle_index(idx), paired, 20, r["start"].tolist(), r["end"].tolist(),
                                   r["chrom"].tolist(), r["mapq"].tolist(), r["flag"].tolist())
    assert counts.tolist() == oc
    assert st[_lib.BS_UNITS] + 1 == os_["total_reads"]
    assert (st[_lib.BS_ASSIGNED], st[_lib.BS_LOWQ], st[_lib.BS_BADCHROM], st[_lib.BS_QCFAIL]) == \
        (os_["assigned"], os_["lowq"], os_["badchrom"], os_["qcfail"])
    assert counts.sum() > 1000
    engine.set_option("bulk_algo", -1)
    engine.set_option("stab_shift", 0)


@pytest.mark.parametrize("algo", [0, 1, 2])
def test_bulk_dense_overlaps_and_gapped_reads(engine, algo):
    """Dense TE pile-ups on a tiny chromosome (sets larger than the register set -> exact path),
    long N-gapped SE reads whose two points fall in different cells."""
    idx = synth.synth_index(3, n_te=6000, n_exon=3000, n_gene=40, n_te_names=300, chrom_len=200_000, n_chrom=2)
    r = synth.synth_bulk_reads(4, idx, 50000, paired=False, edge_frac=0.2)
    engine.set_option("bulk_algo", algo)
    engine.upload_index(idx)
    engine.bulk_begin(False, 20)
    engine.bulk_push(len(r["start"]), r["start"], r["end"], r["chrom"], r["mapq"], r["flag"])
    counts, st = engine.bulk_finish()
    oc, os_ = te_oracle.bulk_count(H.oracle_index(idx), False, 20, r["start"].tolist(), r["end"].tolist(),
                                   r["chrom"].tolist(), r["mapq"].tolist(), r["flag"].tolist())
    assert counts.tolist() == oc and st[_lib.BS_ASSIGNED] == os_["assigned"]
    engine.set_option("bulk_algo", -1)


@pytest.mark.parametrize("all_hot", [0, 1])
@pytest.mark.parametrize("paired", [False, True])
def test_bulk_counter_placement(engine, paired, all_hot):
    """all_hot 1: every ensg counter in shared memory (1024-thread CTAs); 0: the 4096 hottest only."""
    idx = synth.synth_index(9, n_te=40000, n_exon=30000, n_gene=6000, n_te_names=900, chrom_len=4_000_000, n_chrom=3)
    r = synth.synth_bulk_reads(19, idx, 80000, paired=paired, edge_frac=0.02)
    engine.set_option("all_hot", all_hot)
    engine.upload_index(idx)
    engine.bulk_begin(paired, 20)
    engine.bulk_push(len(r["start"]), r["start"], r["end"], r["chrom"], r["mapq"], r["flag"])
    counts, st = engine.bulk_finish()
    engine.set_option("all_hot", 1)
    oc, os_ = te_oracle.bulk_count(H.oracle_index(idx), paired, 20, r["start"].tolist(), r["end"].tolist(),
                                   r["chrom"].tolist(), r["mapq"].tolist(), r["flag"].tolist())
    assert counts.tolist() == oc
    assert (st[_lib.BS_ASSIGNED], st[_lib.BS_LOWQ], st[_lib.BS_BADCHROM], st[_lib.BS_QCFAIL]) == \
        (os_["assigned"], os_["lowq"], os_["badchrom"], os_["qcfail"])


@pytest.mark.parametrize("paired", [False, True])
@pytest.mark.parametrize("bulk_mode,second_mode", [(0, 0), (5, 1), (13, 0), (13, 1), (13, 2), (29, 2), (21, 1), (31, 0), (4, 2), (45, 2), (61, 2), (77, 2), (69, 2)])
def test_bulk_kernel_variants(engine, paired, bulk_mode, second_mode):
    """Every variant of the two-pass kernels (tally by reduction / ballot queue / scan queue, shallow / deep pipeline,
    both register sets of the second pass) against the oracle: dense pile-ups, gapped reads, EDGE cells."""
    idx = synth.synth_index(5, n_te=30000, n_exon=12000, n_gene=900, n_te_names=500, chrom_len=1_500_000, n_chrom=3)
    r = synth.synth_bulk_reads(6, idx, 150_001 if not paired else 150_002, paired=paired, edge_frac=0.05)
    engine.set_option("bulk_mode", bulk_mode)
    engine.set_option("second_mode", second_mode)
    try:
        engine.upload_index(idx)
        engine.bulk_begin(paired, 20)
        engine.bulk_push(len(r["start"]), r["start"], r["end"], r["chrom"], r["mapq"], r["flag"])
        counts, st = engine.bulk_finish()
    finally:
        engine.set_option("bulk_mode", 77)
        engine.set_option("second_mode", 2)
    oc, os_ = te_oracle.bulk_count(H.oracle_index(idx), paired, 20, r["start"].tolist(), r["end"].tolist(),
                                   r["chrom"].tolist(), r["mapq"].tolist(), r["flag"].tolist())
    assert counts.tolist() == oc
    assert (st[_lib.BS_ASSIGNED], st[_lib.BS_LOWQ], st[_lib.BS_BADCHROM], st[_lib.BS_QCFAIL]) == \
        (os_["assigned"], os_["lowq"], os_["badchrom"], os_["qcfail"])


def test_bulk_deep_pileup_overflow_path(engine):
    """> BULK_MAX_DISTINCT distinct ensg under one read: the O(h^2) re-walk path."""
    n = 40
    L = np.full(n, 1000, np.int32) + np.arange(n, dtype=np.int32)
    R = np.full(n, 5000, np.int32)
    from te_counter_b200.index import GlbIndex
    idx = GlbIndex(["1"], np.zeros(n, np.int32), L, R, np.arange(n, dtype=np.int32) % 30,
                   np.full(n, 2, np.uint8), np.zeros(n, np.uint8), ["e%02d" % i for i in range(30)])
    engine.upload_index(idx)
    engine.bulk_begin(False, 0)
    start = np.array([2000, 1010, 900], np.int32)
    end = np.array([2100, 1011, 1001], np.int32)
    z8 = np.zeros(3, np.uint8)
    engine.bulk_push(3, start, end, np.zeros(3, np.uint16), z8 + 60, z8)
    counts, st = engine.bulk_finish()
    oc, _ = te_oracle.bulk_count(H.oracle_index(idx), False, 0, start.tolist(), end.tolist(), [0, 0, 0], [60] * 3, [0] * 3)
    assert counts.tolist() == oc


def test_bulk_linearity_and_chunking(engine):
    """counts(A ++ B) == counts(A) + counts(B); pushing in many small batches == one batch."""
    idx = synth.synth_index(5, n_te=20000, n_exon=5000, n_gene=300, chrom_len=2_000_000, n_chrom=2)
    r = synth.synth_bulk_reads(6, idx, 200000, paired=True)
    engine.upload_index(idx)
    cols = [r[k] for k in ("start", "end", "chrom", "mapq", "flag")]

    def run(slices):
        engine.bulk_begin(True, 20)
        for a, b in slices:
            engine.bulk_push(b - a, *[np.ascontiguousarray(c[a:b]) for c in cols])
        return engine.bulk_finish()

    n = len(cols[0])
    whole, st = run([(0, n)])
    parts, st2 = run([(a, min(n, a + 10000)) for a in range(0, n, 10000)])
    a_, _ = run([(0, 70000)])
    b_, _ = run([(70000, n)])
    assert (whole == parts).all() and (st == st2).all()
    assert (whole == a_ + b_).all()


def test_misuse_errors(engine):
    with pytest.raises(_lib.TecError):
        engine.bulk_push(3, *[np.zeros(4, d) for d in (np.int32, np.int32, np.uint16, np.uint8, np.uint8)]) \
            if engine.bulk_begin(True, 20) is None else None
    e2 = _lib.Engine(0)
    with pytest.raises(_lib.TecError):
        e2.bulk_begin(False, 20)          # no index
    e2.close()


@pytest.mark.parametrize("decoder", ["native", "python"])
@pytest.mark.parametrize("name", ["bulk_pe_rand_b", "bulk_se_rand_b"])
def test_bulk_from_bam_file(monkeypatch, tmp_path, name, decoder):
    """A real BAM file (independent writer) -> libtecbam (native) or te_counter_b200/bam.py + reads.py
    (python) -> pinned batches -> CUDA library -> the reference's TSV bytes."""
    import sys
    import te_counter_b200
    from bam_writer import write_bam
    from oracle.ref_runner import CaptureLog
    case = H.load_case(name)
    recs = [dict(r, name=r.get("name", "r%d" % i)) for i, r in enumerate(case["records"])]
    if any(r["end"] <= r["start"] and not r.get("flag", 0) & 4 for r in recs):
        pytest.skip("case has zero-length alignments a file cannot carry")
    path = str(tmp_path / "x.bam")
    write_bam(path, recs)
    monkeypatch.setitem(sys.modules, "pysam", None)
    monkeypatch.setenv("TEC_BAM_DECODER", decoder)
    mte = te_counter_b200.measureTE("test", case["qual"], device=0)
    mte.bind_genome(H.GOLD + "/" + case["glb"])
    mte.load_genome()
    log = CaptureLog()
    res = (mte.parse_bampe if case["paired"] else mte.parse_bamse)(path, strand=False, log=log)
    assert res == case["expected"]["result"]
    out = tmp_path / "o.tsv"
    mte.save_result_bulk(res, str(out), log=log)
    assert out.read_text() == case["expected"]["tsv"]
