"""Shared test plumbing: golden-case loading, packing through the product's host code, and the
array-level oracle calls (oracle/ is test infrastructure; the product never imports it)."""
import glob
import gzip
import json
import os
import re

import numpy as np

from oracle import te_oracle
from oracle.ref_runner import StubRead
from te_counter_b200 import index as tindex
from te_counter_b200 import reads as treads

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names(kind):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLD, "case_*.json.gz"))):
        name = os.path.basename(p)[5:-8]
        if name.startswith(kind):
            out.append(name)
    return out


def load_case(name):
    with gzip.open(os.path.join(GOLD, "case_%s.json.gz" % name), "rb") as fh:
        return json.loads(fh.read().decode())


_INDEX_CACHE = {}


def load_index(glb):
    if glb not in _INDEX_CACHE:
        _INDEX_CACHE[glb] = tindex.load_glb(os.path.join(GOLD, glb))
    return _INDEX_CACHE[glb]


def stub_reads(recs):
    out = []
    for i, r in enumerate(recs):
        tags = [(t, r[t]) for t in ("CB", "CR", "UB", "UR") if r.get(t) is not None]
        out.append(StubRead(r["chrom"], r["start"], r["end"], r.get("mapq", 60), r.get("flag", 0),
                            r.get("name", "r%d" % i), tags))
    return out


def pack_bulk(case, idx, batch=None):
    recs = stub_reads(case["records"])
    cm = treads.ChromMap(idx.chrom_keys)
    b = treads.Batch(max(2, len(recs) + 2))
    treads.fill_bulk(b, iter(recs), cm, case["paired"], case["qual"])
    n = b.n
    return {k: getattr(b, k)[:n].copy() for k in ("start", "end", "chrom", "mapq", "flag")}


class ListWhitelist(treads.Whitelist):
    def __init__(self, barcodes):
        self.id_to_barcode = sorted(set(barcodes))
        self.barcode_to_id = {bc: i for i, bc in enumerate(self.id_to_barcode)}


def pack_sc(case, idx):
    recs = stub_reads(case["records"])
    cm = treads.ChromMap(idx.chrom_keys)
    wl = ListWhitelist(case["whitelist"])
    b = treads.Batch(max(2, len(recs) + 2), sc=True)
    treads.fill_sc(b, iter(recs), cm, wl, case["qual"])
    n = b.n
    arrs = {k: getattr(b, k)[:n].copy() for k in ("start", "end", "chrom", "mapq", "flag", "cell", "umi")}
    return arrs, wl


def oracle_index(idx):
    return te_oracle.Index(idx.chrom_id, idx.L, idx.R, idx.ensg_id, idx.type_code, idx.strand_code,
                           idx.n_ensg, idx.bucket_size)


def log_number(lines, pattern):
    """First integer (with thousands separators) of the first log line matching `pattern`."""
    for l in lines:
        m = re.search(pattern, l)
        if m:
            return int(m.group(1).replace(",", ""))
    raise KeyError(pattern)


def bulk_expected_stats(exp):
    lg = exp["log"]
    return {"total_reads": exp["total_reads"],
            "assigned": log_number(lg, r"^([\d,]+) Reads were assigned to a gene"),
            "lowq": log_number(lg, r"^([\d,]+) Read quality is too low"),
            "badchrom": log_number(lg, r"^([\d,]+) Reads mapped to an invalid chromosome"),
            "qcfail": log_number(lg, r"^([\d,]+) Reads are QC fails")}


def sc_expected_stats(exp):
    lg = exp["log"]
    return {"total_reads": exp["total_reads"],
            "invalid_barcode": log_number(lg, r"^\s*([\d,]+) invalid barcode reads"),
            "already_seen": log_number(lg, r"^\s*([\d,]+) UMI-CB combinations were seen"),
            "lowq": log_number(lg, r"^\s*([\d,]+) Read quality is too low"),
            "qcfail": log_number(lg, r"^\s*([\d,]+) Reads QC failed"),
            "valid": log_number(lg, r"^\s*([\d,]+) total valid reads"),
            "assigned": log_number(lg, r"^\s*Assigned ([\d,]+) "),
            "raw_barcodes": log_number(lg, r"^\s*Observed ([\d,]+) raw barcodes")}
