"""Single cell on several ranks: two processes (gloo for the host-side collectives, both on GPU 0)
push the two halves of one file, exchange the survivors by cell, finalize with the collective
callback and must reproduce the one-process result (= the oracle) exactly, including the bundle
boundaries, the top-cell choice and the Part-2 held-line rule, which are global over the file."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import helpers as H
from oracle import te_oracle
from te_counter_b200 import _lib, dist as tdist, synth

pytestmark = pytest.mark.gpu

COLS = ("start", "end", "chrom", "mapq", "flag", "cell", "umi")
CASES = [(10_000_000, 50, 20, True), (700, 50, 20, False), (64, 20, 5, True), (5, 200, 1000, False)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data():
    idx = synth.synth_index(11, n_te=30000, n_exon=9000, n_gene=600, chrom_len=3_000_000, n_chrom=3)
    r = synth.synth_sc_reads(12, idx, 30000, n_whitelist=300, n_cells=60, umis_per_cell=40)
    return idx, r


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    idx, r = _data()
    n = len(r["start"])
    lo, hi = tdist.shard_units(n, rank, world)
    eng = _lib.Engine(0)
    eng.upload_index(idx)
    res = []
    for bundle_keys, maxcells, pad, strand in CASES:
        eng.sc_begin(20, strand, 300)
        eng.sc_push(hi - lo, *[np.ascontiguousarray(r[k][lo:hi]) for k in COLS])
        tdist.sc_exchange_by_cell(eng, dev)
        nt, nh = eng.sc_finalize(bundle_keys, maxcells, pad)
        ensg, cell, count, hcell, hcount, st = eng.sc_fetch(nt, nh)
        sel = eng.sc_select(maxcells, nh)
        assert (cell % world == rank).all()                     # triples stay with the owner of the cell
        ensg, cell, count = tdist.sc_gather_triples(ensg, cell, count)
        res.append((ensg.tolist(), cell.tolist(), count.tolist(), hcell.tolist(), hcount.tolist(), st.tolist(), sel.tolist()))
    eng.sc_set_collective(None, 0, 1)
    eng.close()
    q.put((rank, res))
    dist.destroy_process_group()


def test_sc_two_ranks_match_oracle():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    idx, r = _data()
    oidx = H.oracle_index(idx)
    for ci, (bundle_keys, maxcells, pad, strand) in enumerate(CASES):
        out = te_oracle.sc_count(oidx, 20, strand, bundle_keys, maxcells, pad, *[r[k].tolist() for k in COLS])
        for rank in range(world):
            ensg, cell, count, hcell, hcount, st, sel = got[rank][ci]
            assert {(e, c): v for e, c, v in zip(ensg, cell, count)} == out["triples"]
            assert list(zip(hcell, hcount)) == sorted(out["cell_hits"])
            s = out["stats"]
            assert st[_lib.SS_UNITS] + 1 == s["total_reads"]
            for k, f in ((_lib.SS_INVALID_BARCODE, "invalid_barcode"), (_lib.SS_ALREADY_SEEN, "already_seen"),
                         (_lib.SS_LOWQ, "lowq"), (_lib.SS_QCFAIL, "qcfail"), (_lib.SS_VALID, "valid"),
                         (_lib.SS_ASSIGNED, "assigned"), (_lib.SS_RAW_BARCODES, "raw_barcodes"), (_lib.SS_BUNDLES, "n_bundles")):
                assert st[k] == s[f], (f, rank, ci)
            want_sel = sorted(out["cell_hits"], key=lambda t: (-t[1], t[0]))[:maxcells]
            assert sel == [c for c, _ in want_sel]
