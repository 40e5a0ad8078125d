"""
TEST INFRASTRUCTURE -- generates tests/golden/ by running the UNMODIFIED reference
(/root/reference, through oracle/ref_runner.py) on synthetic inputs.  Run in the build container:

    python oracle/make_golden.py

Outputs (committed):
    tests/golden/idx_*.glb        indices pickled by the reference's own genelist.save()
    tests/golden/case_*.json.gz   inputs + what the reference returned / wrote / logged

Each case records `mode`: "vanilla" (stock reference) or "ordered" (the module-global name `set`
bound to an insertion-ordered set -- the canonical reading of te_count.py:452).
Vanilla cases are built so that no (cell, UMI) key is hash-order dependent; the generator asserts
vanilla == ordered on them.
"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_runner as rr      # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
BS = 10000

TOY_FEATURES = [
    dict(chr="chr1", left=1000, right=2000, strand="+", type="protein_coding", ensg="ENSG1"),
    dict(chr="chr1", left=1500, right=1800, strand="-", type="TE", ensg="LINE:L1:L1Md"),
    dict(chr="chr1", left=5000, right=5300, strand="-", type="TE", ensg="SINE:Alu:B1"),
    dict(chr="chr1", left=20000, right=20500, strand="+", type="lncRNA", ensg="ENSG2"),
    dict(chr="chrX", left=9990, right=10010, strand="+", type="pseudogene", ensg="ENSG3"),
]


def random_features(rng, n, chroms, chrom_len, with_snrna=True):
    """Mixed-type features with many on exact bucket edges, zero-length ones, long ones, same-ensg
    on both strands, duplicates (GENCODE exons repeat per transcript)."""
    types = ["protein_coding", "lncRNA", "lincRNA", "TE", "TE", "TE", "pseudogene"]
    if with_snrna:
        types.append("snRNA")
    te_names = ["LINE:L1:L1Md_%d" % i for i in range(6)] + ["SINE:Alu:B%d" % i for i in range(4)] + \
               ["LTR:ERVK:IAP%d" % i for i in range(4)] + ["DNA:hAT:Charlie%d" % i for i in range(3)]
    genes = ["ENSG%05d" % i for i in range(max(4, n // 6))]
    gene_strand = {g: "+-"[int(rng.integers(2))] for g in genes}
    feats = []
    for _ in range(n):
        c = chroms[int(rng.integers(len(chroms)))]
        t = types[int(rng.integers(len(types)))]
        mode = int(rng.integers(10))
        if mode == 0:      # left on a bucket edge
            left = int(rng.integers(0, chrom_len // BS)) * BS
            right = left + int(rng.integers(0, 700))
        elif mode == 1:    # right on / next to a bucket edge
            right = int(rng.integers(1, chrom_len // BS)) * BS - int(rng.integers(0, 2))
            left = max(0, right - int(rng.integers(0, 700)))
        elif mode == 2:    # long, spans several buckets
            left = int(rng.integers(0, chrom_len - 3 * BS))
            right = left + int(rng.integers(BS, 3 * BS))
        elif mode == 3:    # zero / one bp
            left = int(rng.integers(0, chrom_len))
            right = left + int(rng.integers(0, 2))
        else:
            left = int(rng.integers(0, chrom_len - 800))
            right = left + int(rng.integers(20, 800))
        if t == "TE":
            ensg = te_names[int(rng.integers(len(te_names)))]
            strand = "+-"[int(rng.integers(2))]
        elif t == "snRNA":
            ensg = "U%d" % int(rng.integers(3))
            strand = "+-"[int(rng.integers(2))]
        else:
            ensg = genes[int(rng.integers(len(genes)))]
            strand = gene_strand[ensg] if rng.random() < 0.9 else "+-"[int(rng.integers(2))]
        feats.append(dict(chr=c, left=left, right=right, strand=strand, type=t, ensg=ensg))
        if rng.random() < 0.1:
            feats.append(dict(feats[-1]))
    return feats


def edge_positions(rng, feats, n, chrom_len):
    """Read coordinates concentrated on feature edges and bucket edges."""
    pos = []
    for _ in range(n):
        m = int(rng.integers(6))
        if m == 0:
            p = int(rng.integers(0, chrom_len))
        elif m == 1:
            p = int(rng.integers(0, chrom_len // BS + 1)) * BS + int(rng.integers(-2, 3))
        else:
            f = feats[int(rng.integers(len(feats)))]
            p = (f["left"] if rng.random() < 0.5 else f["right"]) + int(rng.integers(-3, 4))
        pos.append(max(0, p))
    return pos


def bulk_records(rng, feats, n, chroms, chrom_len, paired):
    recs = []
    extra_chroms = ["chrUn_KI270", "chr9_alt", "chrEBV", "chrx", "2"]
    for i in range(n):
        f = feats[int(rng.integers(len(feats)))]
        if rng.random() < 0.7:
            chrom = f["chr"]
        elif rng.random() < 0.8:
            chrom = chroms[int(rng.integers(len(chroms)))]
        else:
            chrom = extra_chroms[int(rng.integers(len(extra_chroms)))]
        start = edge_positions(rng, [f], 1, chrom_len)[0]
        m = int(rng.integers(4))
        if m == 0:
            end = start + int(rng.integers(1, 4))
        elif m == 1:
            end = start + 100 + int(rng.integers(0, 3000))
        elif m == 2:
            end = max(start + 1, edge_positions(rng, [f], 1, chrom_len)[0])
        else:
            end = start + int(rng.integers(20, 151))
        flag = 0
        if rng.random() < 0.03:
            flag |= 0x4
        if rng.random() < 0.04:
            flag |= 0x400
        if rng.random() < 0.02:
            flag |= 0x200
        if rng.random() < 0.5:
            flag |= 0x10
        if rng.random() < 0.05:
            flag |= 0x100            # secondary: not filtered by the reference
        r = rng.random()
        mapq = 255 if r < 0.6 else (int(rng.integers(20, 60)) if r < 0.8 else int(rng.integers(0, 20)))
        recs.append(dict(chrom=chrom, start=start, end=end, mapq=mapq, flag=flag, name="q%d" % (i // 2 if paired else i)))
    if rng.random() < 0.5:
        recs[0]["start"] = 0
        recs[0]["end"] = 50
    return recs


def sc_records(rng, feats, n, chroms, chrom_len, barcodes, bad_barcodes, umi_len, n_umi, ambiguous, strand=False, qual=20):
    """ambiguous=False keeps every (cell, UMI) key's chrom:strand sequence in the form a..a[b]
    within the whole file, so the stock reference is hash-seed independent on it."""
    alphabet = "ACGNT"
    umis = ["".join(alphabet[int(x)] for x in rng.integers(0, 5, size=int(rng.integers(max(1, umi_len - 2), umi_len + 1))))
            for _ in range(n_umi)]
    weights = rng.lognormal(0, 1.5, size=len(barcodes))
    weights /= weights.sum()
    recs = []
    key_state = {}        # key -> [first cs, n distinct added]
    extra_chroms = ["chrUn_KI270", "chr9_alt", "chrEBV", "7"]
    for i in range(n):
        f = feats[int(rng.integers(len(feats)))]
        r = rng.random()
        if r < 0.85:
            chrom = f["chr"]
        elif r < 0.93:
            chrom = chroms[int(rng.integers(len(chroms)))]
        else:
            chrom = extra_chroms[int(rng.integers(len(extra_chroms)))]
        start = edge_positions(rng, [f], 1, chrom_len)[0]
        end = start + (int(rng.integers(1, 4)) if rng.random() < 0.2 else int(rng.integers(20, 151)))
        if rng.random() < 0.1:
            end = start + 100 + int(rng.integers(0, 25000))
        flag = 0
        if rng.random() < 0.02:
            flag |= 0x4
        if rng.random() < 0.03:
            flag |= 0x400
        if rng.random() < 0.01:
            flag |= 0x200
        if rng.random() < 0.5:
            flag |= 0x10
        rr_ = rng.random()
        mapq = 255 if rr_ < 0.7 else (int(rng.integers(20, 60)) if rr_ < 0.9 else int(rng.integers(0, 20)))
        if rng.random() < 0.05:
            bc = bad_barcodes[int(rng.integers(len(bad_barcodes)))]
        else:
            bc = barcodes[int(rng.choice(len(barcodes), p=weights))]
        umi = umis[int(rng.integers(len(umis)))]
        rec = dict(chrom=chrom, start=start, end=end, mapq=mapq, flag=flag)
        rec["CB" if rng.random() < 0.8 else "CR"] = bc
        rec["UB" if rng.random() < 0.8 else "UR"] = umi
        if rng.random() < 0.3 and recs:
            # repeat an earlier read's key and place (the ~99 % duplicate case, te_count.py:450)
            prev = recs[int(rng.integers(len(recs)))]
            for k in ("chrom", "CB", "CR", "UB", "UR"):
                rec.pop(k, None)
                if k in prev:
                    rec[k] = prev[k]
            rec["flag"] = (rec["flag"] & ~0x10) | (prev["flag"] & 0x10)
            if rng.random() < 0.7:
                rec["start"], rec["end"] = prev["start"], prev["end"]
        recs.append(rec)
    if not ambiguous:
        out = []
        for rec in recs:
            passes = not (rec["flag"] & 0x604) and rec["mapq"] >= qual
            bc = rec.get("CB", rec.get("CR"))
            ck = rec["chrom"].replace("chr", "")
            if passes and bc in barcodes and not ("_" in ck or "alt" in ck):
                key = (bc, rec.get("UB", rec.get("UR")))
                cs = (ck, (rec["flag"] & 0x10) if strand else None)
                st = key_state.setdefault(key, {"first": cs, "closed": False})
                if st["closed"]:
                    continue                      # drop: a read after the set has 2 members
                if st["first"] != cs:
                    st["closed"] = True           # the single trailing non-a
            out.append(rec)
        recs = out
    return recs


def jsonable(o):
    if isinstance(o, dict):
        return {str(k): jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [jsonable(v) for v in o]
    if isinstance(o, (np.integer,)):
        return int(o)
    return o


def save_case(name, payload):
    p = os.path.join(GOLD, "case_%s.json.gz" % name)
    with gzip.GzipFile(p, "wb", mtime=0) as fh:
        fh.write(json.dumps(jsonable(payload), sort_keys=True).encode())
    print("wrote", p, os.path.getsize(p))


def main():
    os.makedirs(GOLD, exist_ok=True)
    rng = np.random.default_rng(20261018)
    mod = rr.load_reference()

    # ------------------------------------------------------------------ indices
    indices = {"toy": TOY_FEATURES}
    chroms3 = ["chr1", "chr2", "chrX"]
    indices["rand_a"] = random_features(rng, 400, chroms3, 120000)
    indices["rand_b"] = random_features(rng, 1500, ["chr1", "chr7", "chrY", "chrM"], 260000)
    indices["rand_sc"] = random_features(rng, 600, chroms3, 120000, with_snrna=True)
    chrom_len = {"toy": 30000, "rand_a": 120000, "rand_b": 260000, "rand_sc": 120000}
    chrom_names = {"toy": ["chr1", "chrX"], "rand_a": chroms3, "rand_b": ["chr1", "chr7", "chrY", "chrM"],
                   "rand_sc": chroms3}
    for k, feats in indices.items():
        rr.build_index(mod, feats, os.path.join(GOLD, "idx_%s.glb" % k))
        # also keep the reference's own bucket hash for the index-reader test
        gl = mod.miniglbase.glload(os.path.join(GOLD, "idx_%s.glb" % k))
        save_case("index_%s" % k, {"kind": "index", "glb": "idx_%s.glb" % k, "features": feats,
                                   "buckets": {c: {str(b): ids for b, ids in d.items()} for c, d in gl.buckets.items()},
                                   "all_feature_names": sorted(set(gl["ensg"]))})

    # ------------------------------------------------------------------ bulk
    def bulk_case(name, idx, recs, paired, qual=20):
        out = rr.run_bulk(mod, os.path.join(GOLD, "idx_%s.glb" % idx), recs, paired, qual)
        save_case(name, {"kind": "bulk", "glb": "idx_%s.glb" % idx, "paired": paired, "qual": qual,
                         "records": recs, "expected": out, "mode": "vanilla"})

    se = [("chr1", 1600, 1650), ("chr1", 5100, 5150), ("chr1", 999, 1000), ("chr1", 900, 1001),
          ("chr1", 20000, 20050), ("chr1", 19950, 20000), ("chrX", 10000, 10005), ("chr2", 5, 50),
          ("chr1", 1600, 1650, 3), ("chr1", 2000, 2050), ("chr1", 1950, 2001)]
    bulk_case("bulk_se_appendixA", "toy",
              [dict(chrom=r[0], start=r[1], end=r[2], mapq=(r[3] if len(r) > 3 else 60), flag=0) for r in se], False)
    pe = [("p1/1", "chr1", 1600), ("p1/2", "chr1", 1700), ("p2/1", "chr1", 5100), ("p2/2", "chr1", 20100)]
    bulk_case("bulk_pe_appendixA", "toy",
              [dict(name=r[0], chrom=r[1], start=r[2], end=r[2] + 50, mapq=60, flag=0) for r in pe], True)
    bulk_case("bulk_se_empty", "toy", [], False)
    bulk_case("bulk_pe_odd", "toy",
              [dict(chrom="chr1", start=1600, end=1650, mapq=60, flag=0, name="a")] * 3, True)
    for idx in ("rand_a", "rand_b"):
        n = 3000 if idx == "rand_a" else 6000
        recs = bulk_records(rng, indices[idx], n, chrom_names[idx], chrom_len[idx], False)
        bulk_case("bulk_se_%s" % idx, idx, recs, False)
        recs = bulk_records(rng, indices[idx], n + 1, chrom_names[idx], chrom_len[idx], True)
        bulk_case("bulk_pe_%s" % idx, idx, recs, True)
    recs = bulk_records(rng, indices["rand_a"], 1500, chrom_names["rand_a"], chrom_len["rand_a"], False)
    bulk_case("bulk_se_rand_a_q0", "rand_a", recs, False, qual=0)
    bulk_case("bulk_se_rand_a_q30", "rand_a", recs, False, qual=30)

    # ------------------------------------------------------------------ single cell
    def sc_case(name, idx, recs, whitelist, maxcells, strand, bundle_keys=None, pad=None, mode="vanilla", qual=20):
        m = rr.load_reference(bundle_keys=bundle_keys, pad=pad, ordered_sets=(mode == "ordered"))
        glb = os.path.join(GOLD, "idx_%s.glb" % idx)
        out = rr.run_sc(m, glb, recs, whitelist, maxcells, strand=strand, qual=qual)
        if mode == "vanilla":
            m2 = rr.load_reference(bundle_keys=bundle_keys, pad=pad, ordered_sets=True)
            out2 = rr.run_sc(m2, glb, recs, whitelist, maxcells, strand=strand, qual=qual)
            for k in ("result", "barcodes", "barcode_order", "tsv", "freq", "total_reads"):
                assert out[k] == out2[k], (name, k)
            assert [l for l in out["log"] if "tmp" not in l] == [l for l in out2["log"] if "tmp" not in l]
        save_case(name, {"kind": "sc", "glb": "idx_%s.glb" % idx, "strand": strand, "qual": qual,
                         "whitelist": whitelist, "maxcells": maxcells,
                         "bundle_keys": bundle_keys if bundle_keys is not None else 10000000,
                         "pad": pad if pad is not None else 1000,
                         "records": recs, "expected": out, "mode": mode})

    sc = [("chr1", 1600, 1650, "AAAA", "U1"), ("chr1", 1600, 1650, "AAAA", "U2"), ("chr1", 1600, 1650, "AAAA", "U3"),
          ("chr1", 1610, 1660, "AAAA", "U3"), ("chr1", 5100, 5150, "CCCC", "U1"), ("chr1", 5100, 5150, "CCCC", "U2"),
          ("chr1", 20000, 20050, "GGGG", "U9", 0x10), ("chr1", 1600, 1650, "NNNN", "U1"),
          ("chr1", 5100, 5150, "TTTT", "U1"), ("chr1", 5100, 5150, "TTTT", "U2"), ("chrX", 10000, 10005, "TTTT", "U3")]
    # the appendix uses U1/U2..; UMIs here must be nucleotides, map U<n> -> a nucleotide string
    umap = {"U1": "AAC", "U2": "AAG", "U3": "AAT", "U9": "TTT"}
    recs = [dict(chrom=r[0], start=r[1], end=r[2], CB=r[3], UB=umap[r[4]], mapq=60, flag=(r[5] if len(r) > 5 else 0)) for r in sc]
    wl4 = ["AAAA", "CCCC", "GGGG", "TTTT"]
    sc_case("sc_appendixA", "toy", recs, wl4, 3, False)
    sc_case("sc_appendixA_strand", "toy", recs, wl4, 3, True)

    bases = "ACGT"
    def make_barcodes(n, k):
        s = set()
        while len(s) < n:
            s.add("".join(bases[int(x)] for x in rng.integers(0, 4, size=k)))
        return sorted(s)

    feats = indices["rand_sc"]
    bcs = make_barcodes(40, 8)
    wl, bad = bcs[:32], bcs[32:]
    for strand in (False, True):
        tag = "_strand" if strand else ""
        recs = sc_records(rng, feats, 4000, chrom_names["rand_sc"], chrom_len["rand_sc"], wl, bad, 5, 60, ambiguous=False, strand=strand)
        sc_case("sc_rand_det%s" % tag, "rand_sc", recs, wl + [""], 5, strand, pad=3)
        sc_case("sc_rand_det_bundles%s" % tag, "rand_sc", recs, wl, 4, strand, bundle_keys=97, pad=2)
        recs = sc_records(rng, feats, 4000, chrom_names["rand_sc"], chrom_len["rand_sc"], wl, bad, 4, 25, ambiguous=True)
        sc_case("sc_rand_amb%s" % tag, "rand_sc", recs, wl, 6, strand, pad=4, mode="ordered")
        sc_case("sc_rand_amb_bundles%s" % tag, "rand_sc", recs, wl, 6, strand, bundle_keys=53, pad=1, mode="ordered")
        sc_case("sc_rand_amb_tinybundles%s" % tag, "rand_sc", recs[:800], wl, 3, strand, bundle_keys=3, pad=0, mode="ordered")
    # stock pad (+1000) with more cells than maxcells+1000, multi bundle
    bcs = make_barcodes(1300, 8)
    wl, bad = bcs[:1280], bcs[1280:]
    recs = sc_records(rng, feats, 9000, chrom_names["rand_sc"], chrom_len["rand_sc"], wl, bad, 4, 40, ambiguous=False, strand=True)
    sc_case("sc_rand_det_stockpad", "rand_sc", recs, wl, 20, True, bundle_keys=1500)
    recs = sc_records(rng, feats, 2000, chrom_names["rand_sc"], chrom_len["rand_sc"], wl, bad, 4, 40, ambiguous=False, strand=False, qual=0)
    sc_case("sc_rand_det_q0", "rand_sc", recs, wl, 2000, False, qual=0)
    rr.cleanup()


if __name__ == "__main__":
    main()
