#!/usr/bin/env python3
"""A/B of bulk engine options on the BASELINE configs[3] workload, reads generated once:
    python tools/bulk_sweep.py --workload bulk_pe --configs "stab_shift=10,bulk_mode=3;stab_shift=11,bulk_mode=3"
Prints one JSON line per configuration (device-timed ms per pass, deferred / slow units, counts checksum)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="bulk_pe")
    ap.add_argument("--records", type=int, default=500_000_000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--configs", default="bulk_algo=2")
    ap.add_argument("--shard", default="0,1", help="rank,world: reads restricted to that slice of the genome (how much of the table is touched)")
    a = ap.parse_args()
    import torch
    from te_counter_b200 import _lib, synth
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    paired = a.workload == "bulk_pe"
    idx = synth.synth_index()
    reads = synth.synth_bulk_reads(synth.SEED, idx, a.records, paired=paired, device=dev, as_numpy=False,
                                   shard=tuple(int(x) for x in a.shard.split(",")))
    ptrs = [reads[k].data_ptr() for k in ("start", "end", "chrom", "mapq", "flag")]
    torch.cuda.synchronize()
    ref = None
    for cfg in a.configs.split(";"):
        eng = _lib.Engine(0)
        for kv in cfg.split(","):
            if kv:
                k, v = kv.split("=")
                eng.set_option(k, int(v))
        eng.upload_index(idx)
        ext = torch.cuda.ExternalStream(eng.stream, device=dev)
        for _ in range(3):
            eng.bulk_begin(paired, 20)
            eng.bulk_push_dev(a.records, *ptrs)
        eng.sync()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
        for e0, e1 in ev:
            eng.bulk_begin(paired, 20)
            e0.record(ext)
            eng.bulk_push_dev(a.records, *ptrs)
            e1.record(ext)
        eng.sync()
        torch.cuda.synchronize()
        ms = [e0.elapsed_time(e1) for e0, e1 in ev]
        counts, st = eng.bulk_finish()
        chk = int((counts.astype(np.uint64) * (np.arange(len(counts), dtype=np.uint64) * np.uint64(2654435761) + np.uint64(1))).sum() & np.uint64(0xFFFFFFFFFFFF))
        if ref is None:
            ref = (chk, st[:5].tolist())
        out = {"config": cfg, "shard": a.shard, "workload": a.workload, "records": a.records, "ms_mean": float(np.mean(ms)), "ms_min": float(np.min(ms)),
               "records_per_s": a.records / (float(np.mean(ms)) / 1e3), "deferred": eng.get_info("last_deferred_units"),
               "slow": eng.get_info("last_slow_units"), "table_bytes": eng.get_info("stab_bytes"), "checksum": chk,
               "same_as_first": (chk, st[:5].tolist()) == ref, "stats": [int(x) for x in st[:5]]}
        print(json.dumps(out), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
