"""N > 1 host logic on CPU: two gloo ranks each count their shard of one bulk workload (oracle
engine standing in for the GPU), merge with dist.allreduce_counts_host, and must reproduce the
single-rank result (bulk shards with no data-path collective besides the merge)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

import helpers as H
from oracle_engine import OracleEngine
from te_counter_b200 import _lib, dist as tdist, synth

COLS = ("start", "end", "chrom", "mapq", "flag")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _count(idx, r, lo, hi, paired):
    eng = OracleEngine(0)
    eng.upload_index(idx)
    eng.bulk_begin(paired, 20)
    eng.bulk_push(hi - lo, *[r[k][lo:hi] for k in COLS])
    return eng.bulk_finish()


def _worker(rank, world, port, paired, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = synth.synth_index(21, n_te=3000, n_exon=900, n_gene=60, chrom_len=400_000, n_chrom=2)
    r = synth.synth_bulk_reads(22, idx, 6000, paired=paired)
    n_units = 3000 if paired else 6000
    lo, hi = tdist.shard_units(n_units, rank, world)
    k = 2 if paired else 1
    counts, st = _count(idx, r, lo * k, hi * k, paired)
    counts, st = tdist.allreduce_counts_host(counts, st)
    if rank == 0:
        q.put((counts.tolist(), st.tolist()))
    dist.destroy_process_group()


def _run(paired):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, paired, q)) for r in range(world)]
    for p in ps:
        p.start()
    got = q.get(timeout=120)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    idx = synth.synth_index(21, n_te=3000, n_exon=900, n_gene=60, chrom_len=400_000, n_chrom=2)
    r = synth.synth_bulk_reads(22, idx, 6000, paired=paired)
    counts, st = _count(idx, r, 0, 6000, paired)
    assert got[0] == counts.tolist()
    # units / assigned / lowq / badchrom / qcfail add up across ranks
    assert got[1][:5] == st.tolist()[:5]


def test_two_rank_merge_se():
    _run(False)


def test_two_rank_merge_pe():
    _run(True)


def test_shard_units_cover():
    for n in (0, 1, 7, 1000, 12345):
        for w in (1, 2, 3, 8):
            spans = [tdist.shard_units(n, r, w, align=2) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(lo % 2 == 0 for lo, _ in spans)
