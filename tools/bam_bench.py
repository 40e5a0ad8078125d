"""Throughput of the host BAM decoder (libtecbam, include/tecbam.h): records/s from a BAM file on
disk (page cache warm) to filled structure-of-arrays batches, per thread count, with the Python
packing it replaces (bam.py + reads.fill_*) timed on a bounded sample beside it.  Prints one JSON
line.  No GPU involved: the batches are the arrays tec_bulk_push / tec_sc_push take.

    python tools/bam_bench.py file.bam --mode sc --whitelist wl.txt --threads 1,4,16,64
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from te_counter_b200 import bam, fastbam, reads             # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("bam")
    ap.add_argument("--mode", choices=["se", "pe", "sc"], default="pe")
    ap.add_argument("--whitelist")
    ap.add_argument("--threads", default="1,2,4,8")
    ap.add_argument("--python-records", type=int, default=200000)
    ap.add_argument("--batch", type=int, default=1 << 20)
    a = ap.parse_args()
    chrom_keys = [str(c) for c in list(range(1, 23)) + ["X", "Y", "M"]]
    wl = reads.Whitelist(a.whitelist) if a.mode == "sc" else None
    size = os.path.getsize(a.bam)
    with open(a.bam, "rb") as fh:                           # warm the page cache
        while fh.read(1 << 24):
            pass
    out = {"file_bytes": size, "mode": a.mode, "batch_records": a.batch, "native": [], "host_cores": os.cpu_count()}
    ref = None
    for t in [int(x) for x in a.threads.split(",")]:
        f = fastbam.NativeBam(a.bam, threads=t)
        f.bind(reads.ChromMap(chrom_keys), wl)
        b = reads.Batch(a.batch, sc=a.mode == "sc")
        t0 = time.perf_counter()
        n, more, chk = 0, True, 0
        while more:
            more = f.fill_sc(b, 20) if a.mode == "sc" else f.fill_bulk(b, a.mode == "pe", 20)
            n += b.n
            chk += int(b.start[:b.n].astype(np.int64).sum()) + int(b.end[:b.n].astype(np.int64).sum())
        dt = time.perf_counter() - t0
        c = f.counters()
        f.close()
        if ref is None:
            ref = chk
        assert chk == ref, "thread counts disagree"
        out["native"].append({"threads": c["threads"], "records_per_s": n / dt, "compressed_GBps": size / dt / 1e9,
                              "uncompressed_GBps": c["uncompressed_bytes"] / dt / 1e9, "seconds": dt})
        out["records"] = n
    if a.python_records:
        f = bam.AlignmentFile(a.bam, "r")
        cm = reads.ChromMap(chrom_keys)
        b = reads.Batch(a.python_records - (a.python_records & 1), sc=a.mode == "sc")
        t0 = time.perf_counter()
        if a.mode == "sc":
            reads.fill_sc(b, f, cm, wl, 20)
        else:
            reads.fill_bulk(b, f, cm, a.mode == "pe", 20)
        dt = time.perf_counter() - t0
        f.close()
        out["python_packing"] = {"records_per_s": b.n / dt, "sample_records": b.n, "cores": 1,
                                 "what": "te_counter_b200/bam.py + reads.fill_* (stand-in for the pysam loop)"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
