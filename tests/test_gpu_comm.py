"""Collectives issued by the library (include/tecount.h: tec_comm_*, tec_bulk_allreduce, tec_sc_exchange,
tec_sc_allgather_triples; csrc/collective.cuh, sc_comm.cuh).

On one GPU the communicator has one rank: every call runs its real NCCL path (all-gather of the counts, grouped
send / receive to itself, import, the sort of the gathered triples) and the results must be the oracle's.  With two
or more GPUs on the box the two-rank parity run of tools/sc_dist_parity.py (bit-exact against the C++ oracle over
the whole file) is launched under torchrun as well."""
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H
from oracle import te_oracle
from te_counter_b200 import _lib, synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ("start", "end", "chrom", "mapq", "flag", "cell", "umi")


@pytest.fixture(scope="module")
def engine():
    eng = _lib.Engine(0)
    try:
        eng.comm_init(eng.comm_unique_id(), 0, 1)
    except _lib.TecError as e:
        eng.close()
        pytest.skip("NCCL not available: %s" % e)
    yield eng
    eng.close()


def test_comm_info_and_bad_arguments(engine):
    assert engine.comm_world() == (0, 1)
    with pytest.raises(ValueError):
        engine.comm_init(b"short", 0, 1)
    with pytest.raises(_lib.TecError):
        engine.comm_init(b"\0" * _lib.COMM_ID_BYTES, 3, 2)           # rank outside the world


def test_bulk_allreduce_single_rank_is_identity(engine):
    idx = synth.synth_index(21, n_te=20000, n_exon=6000, n_gene=400, chrom_len=2_000_000, n_chrom=3)
    oidx = H.oracle_index(idx)
    engine.upload_index(idx)
    r = synth.synth_bulk_reads(22, idx, 30000, paired=True, edge_frac=0.05)
    engine.bulk_begin(True, 20)
    engine.bulk_push(len(r["start"]), r["start"], r["end"], r["chrom"], r["mapq"], r["flag"])
    engine.bulk_allreduce()
    counts, st = engine.bulk_finish()
    oc, os_ = te_oracle.bulk_count(oidx, True, 20, *[r[k].tolist() for k in ("start", "end", "chrom", "mapq", "flag")])
    assert counts.tolist() == oc
    assert int(st[_lib.BS_ASSIGNED]) == os_["assigned"]


@pytest.mark.parametrize("bundle_keys,maxcells,pad,strand", [(10_000_000, 50, 20, True), (64, 20, 5, False)])
def test_sc_exchange_and_allgather_single_rank(engine, bundle_keys, maxcells, pad, strand):
    idx = synth.synth_index(11, n_te=30000, n_exon=9000, n_gene=600, chrom_len=3_000_000, n_chrom=3)
    r = synth.synth_sc_reads(12, idx, 30000, n_whitelist=300, n_cells=60, umis_per_cell=40)
    oidx = H.oracle_index(idx)
    engine.upload_index(idx)
    engine.sc_begin(20, strand, 300)
    engine.sc_push(len(r["start"]), *[r[k] for k in COLS])
    n_surv = engine.sc_survivors()
    assert engine.sc_exchange() == n_surv                              # one rank owns every cell
    nt, nh = engine.sc_finalize(bundle_keys, maxcells, pad)
    assert engine.sc_allgather_triples() == nt
    ensg, cell, count, hcell, hcount, st = engine.sc_fetch(nt, nh)
    out = te_oracle.sc_count(oidx, 20, strand, bundle_keys, maxcells, pad, *[r[k].tolist() for k in COLS])
    assert {(int(e), int(c)): int(v) for e, c, v in zip(ensg, cell, count)} == out["triples"]
    key = ensg.astype(np.int64) << 32 | cell.astype(np.int64)
    assert (np.diff(key) > 0).all()                                    # ascending in (ensg, cell)
    assert list(zip(hcell.tolist(), hcount.tolist())) == sorted(out["cell_hits"])
    assert int(st[_lib.SS_VALID]) == out["stats"]["valid"] and int(st[_lib.SS_BUNDLES]) == out["stats"]["n_bundles"]


def test_two_ranks_over_nccl_match_the_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (the driver's scaling run and tools/gpu_passes/gpu_round2_n2.sh cover it)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(ROOT, "tools", "sc_dist_parity.py"), "--records-per-rank", "3000000",
           "--bundle-keys", "400000"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert '"bit_exact": true' in p.stdout and "library (NCCL" in p.stdout
