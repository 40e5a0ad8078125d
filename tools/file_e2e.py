"""File-to-result run of measureTE on a synthetic BAM: BGZF file on disk -> libtecbam (host threads)
-> pinned batches -> libtecount (GPU) -> counts, timed by wall clock around the public call
(parse_bampe / parse_bamse / sc_parse_bamse).  The counts are then checked against the C / C++
oracle fed with the arrays the decoder produced (decode parity itself: tests/test_fastbam.py).
Prints one JSON line.  Needs a GPU.

    python tools/file_e2e.py --mode pe --records 8000000
"""
import argparse
import json
import logging
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import te_counter_b200                                              # noqa: E402
from te_counter_b200 import fastbam, reads, synth                   # noqa: E402


def decode_all(path, mode, chrom_keys, wl):
    f = fastbam.NativeBam(path)
    f.bind(reads.ChromMap(chrom_keys), wl)
    cols = {k: [] for k in ("start", "end", "chrom", "mapq", "flag") + (("cell", "umi") if mode == "sc" else ())}
    b = reads.Batch(1 << 20, sc=mode == "sc")
    t0 = time.perf_counter()
    more = True
    while more:
        more = f.fill_sc(b, 20) if mode == "sc" else f.fill_bulk(b, mode == "pe", 20)
        for k in cols:
            cols[k].append(getattr(b, k)[:b.n].copy())
    dt = time.perf_counter() - t0
    threads = f.counters()["threads"]
    f.close()
    return {k: np.concatenate(v) for k, v in cols.items()}, dt, threads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["se", "pe", "sc"], default="pe")
    ap.add_argument("--records", type=int, default=8000000)
    ap.add_argument("--dir", default="/tmp")
    ap.add_argument("--maxcells", type=int, default=10000)
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value")
    ap.add_argument("--decoder", choices=["native", "gpu"], default="native", help="TEC_BAM_DECODER for the timed call")
    a = ap.parse_args()
    path = os.path.join(a.dir, "tec_synth_%s.bam" % a.mode)
    wlf = os.path.join(a.dir, "tec_synth_wl.txt")
    t0 = time.perf_counter()
    subprocess.run([sys.executable, "-m", "te_counter_b200.synth_bam", path, "--records", str(a.records),
                    "--mode", a.mode, "--whitelist", wlf], check=True, stdout=subprocess.DEVNULL, cwd=ROOT)
    t_make = time.perf_counter() - t0
    idx = synth.synth_index()
    log = logging.getLogger("file_e2e")
    log.addHandler(logging.NullHandler())
    log.propagate = False
    mte = te_counter_b200.measureTE("file_e2e", 20)
    mte.genome = idx
    mte.all_feature_names = idx.names
    mte.load_genome = lambda: None                      # the synthetic index has no .glb file
    for kv in a.opt:
        k, v = kv.split("=")
        mte._engine().set_option(k, int(v))
    mte._engine()                                       # index upload + table build: outside the timed call, as load_genome is
    os.environ["TEC_BAM_DECODER"] = a.decoder
    out = {"mode": a.mode, "decoder": a.decoder, "file_bytes": os.path.getsize(path), "host_cores": os.cpu_count(), "make_bam_s": t_make}
    runs = []
    for _ in range(3):
        t0 = time.perf_counter()
        if a.mode == "sc":
            res = mte.sc_parse_bamse(path, whitelistfilename=wlf, strand=True, log=log, label="x", maxcells=a.maxcells)
        else:
            res = (mte.parse_bampe if a.mode == "pe" else mte.parse_bamse)(path, log=log)
        runs.append(time.perf_counter() - t0)
    n = mte.total_reads - 1
    n_rec = n * 2 if a.mode == "pe" else n
    out.update({"records": n_rec, "file_to_result_s": runs, "file_to_result_records_per_s": n_rec / min(runs)})
    if getattr(mte, "_bam_info", None):
        out["device_decoder_last_run"] = mte._bam_info
    wl = reads.Whitelist(wlf) if a.mode == "sc" else None
    cols, dt, threads = decode_all(path, a.mode, idx.chrom_keys, wl)
    out.update({"decode_only_records_per_s": len(cols["start"]) / dt, "decode_threads": threads})
    if not a.no_parity:
        from oracle import te_oracle_c
        if a.mode == "sc":
            o = te_oracle_c.sc_count((idx.chrom_id, idx.L, idx.R, idx.ensg_id, idx.type_code, idx.strand_code), idx.n_chrom,
                                     idx.bucket_size, 20, True, 10000000, a.maxcells, 1000,
                                     *[cols[k] for k in ("start", "end", "chrom", "mapq", "flag", "cell", "umi")])
            o_ensg, o_cell, o_count = o["triples_arrays"]
            ok = (len(o_ensg) == len(res.ensg) and (o_ensg == res.ensg).all() and (o_cell == res.cell).all()
                  and (o_count == res.count).all())
            out["parity"] = {"checker": "oracle/te_oracle_sc.cpp on the decoded arrays", "triples": int(len(o_ensg)),
                             "bundles": o["stats"]["n_bundles"], "equal": bool(ok)}
        else:
            cidx = te_oracle_c.Index(idx.chrom_id, idx.L, idx.R, idx.ensg_id, idx.type_code, idx.n_chrom, idx.n_ensg,
                                     idx.bucket_size)
            oc, _ = te_oracle_c.bulk_count(cidx, a.mode == "pe", 20, cols["start"], cols["end"], cols["chrom"], cols["mapq"],
                                           cols["flag"])
            mine = np.array([res[k] for k in idx.names], dtype=np.int64)
            out["parity"] = {"checker": "oracle/te_oracle_c.c on the decoded arrays", "counted": int(mine.sum()),
                             "equal": bool(np.array_equal(mine, oc))}
        assert out["parity"]["equal"], "file-to-result counts differ from the oracle"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
