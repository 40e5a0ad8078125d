#!/usr/bin/env python3
"""Multi-GPU single-cell parity at scale (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
        --master-port 29555 tools/sc_dist_parity.py --records-per-rank 20000000

Every rank pushes its slice of one coordinate-sorted synthetic file, the survivors are exchanged by
cell over NCCL, tec_sc_finalize runs with the collective callback and real 1e7-key bundles; rank 0
then runs the C++ oracle (oracle/te_oracle_sc.cpp) on the WHOLE file and compares triples, hit cells,
statistics and the selected cells bit for bit.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from te_counter_b200 import _lib, synth, dist as tdist
    from oracle import te_oracle_c
    ap = argparse.ArgumentParser()
    ap.add_argument("--records-per-rank", type=int, default=20_000_000)
    ap.add_argument("--bundle-keys", type=int, default=10_000_000)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    idx = synth.synth_index()
    n_wl, maxcells, pad, strand = 100_000, 10_000, 1000, True
    cols = ("start", "end", "chrom", "mapq", "flag", "cell", "umi")
    r = synth.synth_sc_reads(synth.SEED, idx, args.records_per_rank, n_whitelist=n_wl, device=dev, as_numpy=True, part=(rank, world))
    eng = _lib.Engine(local)
    eng.upload_index(idx)
    in_library = tdist.comm_init(eng) and not os.environ.get("TEC_DIST_CALLBACKS")
    if not in_library:
        eng.comm_destroy()
    eng.sc_begin(20, strand, n_wl)
    n = len(r["start"])
    for a in range(0, n, 1 << 24):
        eng.sc_push(min(n, a + (1 << 24)) - a, *[np.ascontiguousarray(r[k][a:a + (1 << 24)]) for k in cols])
    t0 = time.perf_counter()
    tdist.sc_exchange_by_cell(eng, dev)
    nt, nh = eng.sc_finalize(args.bundle_keys, maxcells, pad)
    t1 = time.perf_counter()
    if in_library:
        nt = eng.sc_allgather_triples()            # NCCL inside the library: every rank holds the job's triples
    ensg, cell, count, hcell, hcount, st = eng.sc_fetch(nt, nh)
    sel = eng.sc_select(maxcells, nh)
    if not in_library:
        ensg, cell, count = tdist.sc_gather_triples(ensg, cell, count)
    parts = [None] * world
    dist.all_gather_object(parts, {k: r[k] for k in cols})
    if rank == 0:
        whole = {k: np.concatenate([p[k] for p in parts]) for k in cols}
        ta = time.perf_counter()
        out = te_oracle_c.sc_count((idx.chrom_id, idx.L, idx.R, idx.ensg_id, idx.type_code, idx.strand_code), idx.n_chrom,
                                   idx.bucket_size, 20, strand, args.bundle_keys, maxcells, pad, *[whole[k] for k in cols])
        tb = time.perf_counter()
        o_ensg, o_cell, o_count = out["triples_arrays"]
        want_sel = [c for c, _ in sorted(out["cell_hits"], key=lambda t: (-t[1], t[0]))[:maxcells]]
        ok = (len(o_ensg) == len(ensg) and (o_ensg == ensg).all() and (o_cell == cell).all() and (o_count == count).all()
              and sorted(out["cell_hits"]) == list(zip(hcell.tolist(), hcount.tolist())) and sel.tolist() == want_sel
              and all(int(st[k]) == out["stats"][f] for k, f in (
                  (_lib.SS_INVALID_BARCODE, "invalid_barcode"), (_lib.SS_ALREADY_SEEN, "already_seen"), (_lib.SS_LOWQ, "lowq"),
                  (_lib.SS_QCFAIL, "qcfail"), (_lib.SS_VALID, "valid"), (_lib.SS_ASSIGNED, "assigned"),
                  (_lib.SS_RAW_BARCODES, "raw_barcodes"), (_lib.SS_BUNDLES, "n_bundles")))
              and int(st[_lib.SS_UNITS]) + 1 == out["stats"]["total_reads"])
        print(json.dumps({"check": "multi-GPU single cell vs C++ oracle", "n_gpus": world, "records_total": int(n * world),
                          "bundles": out["stats"]["n_bundles"], "triples": int(len(o_ensg)), "hit_cells": int(nh),
                          "bit_exact": bool(ok), "collectives": "library (NCCL: tec_sc_exchange, all-reduces inside tec_sc_finalize, tec_sc_allgather_triples)" if in_library else "torch.distributed + callback",
                          "gpu_exchange_plus_finalize_s": t1 - t0, "oracle_s": tb - ta}))
        assert ok
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
