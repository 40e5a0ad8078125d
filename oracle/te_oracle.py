"""
TEST INFRASTRUCTURE -- CPU restatement (oracle) of te_counter's counting hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker.  The product path (te_counter_b200/) never imports it and
fails loudly when the CUDA library is missing.

Parity pin: this restatement is checked against outputs of the UNMODIFIED reference run in the
build container (oracle/ref_runner.py + oracle/make_golden.py -> tests/golden/*.json), including
the SURVEY.md Appendix-A known-answer vectors.  The reference ships no value-pinning tests of its
own for this path, so those generated vectors are the pin.

oracle/te_oracle_c.c is a C restatement of the bulk loop only (same literal algorithm, multi-threaded)
used to check full-size results in seconds; tests/test_oracle_c.py checks it against this module and
against the golden vectors.

Everything follows the reference literally (10 kb bucket hash, Python sets/dicts, the held-line
bundle scan), NOT the closed forms the CUDA kernels use, so that the two are independent.
All citations are relative to /root/reference/.

Array-level inputs (same as the C ABI in include/tecount.h):
  index:  per feature  chrom_id, L, R, ensg_id, type_code, strand_code   (+ n_ensg, bucket_size)
  reads:  start, end, chrom (u16 id; >= n_index_chrom means "not in index"), mapq, flag bits
          (+ sc: cell id (0xFFFFFFFF = not whitelisted), umi code (order preserving u64))
"""
from collections import defaultdict
from operator import itemgetter

# type codes (te_count.py:134-146 branches)
T_OTHER, T_GENE, T_TE, T_SNRNA, T_ENH = 0, 1, 2, 3, 4
# read flag bits of the SoA record
F_UNMAPPED, F_DUP, F_QCFAIL, F_REVERSE, F_NAME_MISMATCH = 1, 2, 4, 8, 16
CHROM_SC_SKIP = 0xFFFE      # '_' or 'alt' in the name (te_count.py:432)
CHROM_INVALID = 0xFFFF      # generic "not a chromosome we know"
CELL_INVALID = 0xFFFFFFFF
# sc strand codes for the cs key: '+', '-', 'NA'
S_PLUS, S_MINUS, S_NA = 0, 1, 2
# feature strand codes: 0 '+', 1 '-', 2/3 any other strings, 255 = the feature dict has no
# 'strand' key (enhancer indices, genome/make.py:129-171) -> the sc path raises KeyError at :661
STRAND_MISSING = 255


class ReferenceCrash(Exception):
    """The reference raises at this input; `kind` is the exception class name it raises."""

    def __init__(self, kind, msg=""):
        Exception.__init__(self, "%s %s" % (kind, msg))
        self.kind = kind


class Index:
    """Flat view of genelist.linearData + the bucket hash of miniglbase/genelist.py:367-380."""

    def __init__(self, chrom_id, L, R, ensg_id, type_code, strand_code, n_ensg, bucket_size=10000):
        self.chrom_id = [int(x) for x in chrom_id]
        self.L = [int(x) for x in L]
        self.R = [int(x) for x in R]
        self.ensg_id = [int(x) for x in ensg_id]
        self.type_code = [int(x) for x in type_code]
        self.strand_code = [int(x) for x in strand_code]
        self.n_ensg = int(n_ensg)
        self.bs = int(bucket_size)
        self.buckets = build_buckets(self.chrom_id, self.L, self.R, self.bs)


def build_buckets(chrom_id, L, R, bs):
    """miniglbase/genelist.py:367-380 -- feature n goes into every bucket of
    range((L//bs)*bs, ((R+bs)//bs)*bs, bs)."""
    buckets = {}
    for n in range(len(L)):
        c = chrom_id[n]
        if c not in buckets:
            buckets[c] = {}
        left_buck = (L[n] // bs) * bs
        right_buck = ((R[n] + bs) // bs) * bs
        for b in range(left_buck, right_buck, bs):
            if b not in buckets[c]:
                buckets[c][b] = []
            buckets[c][b].append(n)
    return buckets


def _bulk_tally(idx, result, counts):
    """te_count.py:128-149 / :243-261.  `result` is the list of hit feature indices."""
    types = set(idx.type_code[i] for i in result)
    ensgs = set(idx.ensg_id[i] for i in result)
    if T_GENE in types:
        for e in ensgs:          # `if ':' in ensgs` (:136) is never true -- dead code
            counts[e] += 1
    elif T_TE in types:
        for e in ensgs:
            counts[e] += 1
    elif T_SNRNA in types:
        for e in ensgs:
            counts[e] += 1
    elif T_ENH in types:
        raise ReferenceCrash("NameError", "barcode (te_count.py:147)")


def bulk_count(idx, paired, qual, start, end, chrom, mapq, flag, candidate=None):
    """te_count.py:42-165 (paired) / :167-277 (single end).

    Returns (counts list[n_ensg], stats dict) where stats['total_reads'] is the reference's
    off-by-one `idx` (te_count.py:77,163).

    candidate: None for the reference's loop.  oracle/te_oracle_ext.py passes a predicate
    (feature index, index of the unit's first record) -> bool to restate the opt-in extensions, which
    are outside the parity claim."""
    counts = [0] * idx.n_ensg
    bs = idx.bs
    assigned = lowq = badchrom = qcfail = 0
    n = len(start)
    pos = 0
    units = 0
    while True:
        units += 1                                   # :77 / :202  idx += 1 before next()
        if paired:
            if pos + 2 > n:                          # StopIteration on 1st or 2nd next()
                break
            r1, r2 = pos, pos + 1
            pos += 2
            if flag[r1] & (F_UNMAPPED | F_DUP | F_QCFAIL):      # :81
                qcfail += 1
                continue
            if flag[r2] & (F_UNMAPPED | F_DUP | F_QCFAIL):      # :84
                qcfail += 1
                continue
            if mapq[r1] < qual:                                  # :88
                lowq += 1
                continue
            if flag[r1] & F_NAME_MISMATCH:                       # :92-94  sys.quit()
                raise ReferenceCrash("AttributeError", "sys.quit (te_count.py:94)")
            c = chrom[r1]                                        # :96
            loc1 = start[r1]                                     # :97
            loc2 = start[r2]                                     # :98
        else:
            if pos + 1 > n:
                break
            r1 = pos
            pos += 1
            if flag[r1] & (F_UNMAPPED | F_DUP | F_QCFAIL):      # :204
                qcfail += 1
                continue
            if mapq[r1] < qual:                                  # :208
                lowq += 1
                continue
            c = chrom[r1]
            loc1 = start[r1]                                     # :213
            loc2 = end[r1]                                       # :214
        if c not in idx.buckets:                                 # :100 / :216
            badchrom += 1
            continue
        left_buck = ((loc1 - 1) // bs) * bs                      # :106
        right_buck = ((loc2 + 1) // bs) * bs                     # :107
        loc_ids = set()
        for buck in (left_buck, right_buck):                     # only those two buckets
            if buck in idx.buckets[c]:
                loc_ids.update(idx.buckets[c][buck])
        result = []
        for i in loc_ids:
            if candidate is not None and not candidate(i, r1):
                continue
            if loc1 >= idx.L[i] and loc1 + 1 <= idx.R[i]:        # :122
                result.append(i)
            if loc2 - 1 >= idx.L[i] and loc2 <= idx.R[i]:        # :125
                result.append(i)
        if result:
            _bulk_tally(idx, result, counts)
            assigned += 1                                        # :149
    return counts, {"total_reads": units, "assigned": assigned, "lowq": lowq,
                    "badchrom": badchrom, "qcfail": qcfail}


# --------------------------------------------------------------------------------- single cell
def sc_count(idx, qual, strand, bundle_keys, maxcells, pad, start, end, chrom, mapq, flag, cell, umi):
    """te_count.py:298-707 with the canonical first-inserted rule of SURVEY.md 8a-9
    (what the reference computes when its sets iterate in insertion order).

    strand: bool (--strand).  bundle_keys: the 1e7 of te_count.py:377.  pad: the +1000 of :502.
    Returns dict(triples={(ensg_id, cell): n}, cell_hits=[(cell, hits) in insertion order],
                 stats={...})"""
    bs = idx.bs
    n = len(start)
    lowq = qcfail = invalid_barcode = already_seen = 0
    barcodes = {}                 # cell id -> raw count, insertion ordered (self.barcodes)
    bundles = []                  # each: sorted list of (cell, umi, [frag...]) == one .bun file
    umis = {}                     # (cell, umi) -> insertion-ordered dict of frag tuples
    units = 0
    pos = 0

    def save_bundle(d):           # :357-368  sorted by (cell id, umi string)
        bundles.append([(k[0], k[1], list(d[k].keys())) for k in sorted(d)])

    while True:
        units += 1                                               # :373
        if len(umis) >= bundle_keys:                             # :377
            save_bundle(umis)
            umis = {}
        if pos >= n:                                             # :393 StopIteration
            break
        r = pos
        pos += 1
        if flag[r] & (F_UNMAPPED | F_DUP | F_QCFAIL):           # :394
            qcfail += 1
            continue
        if mapq[r] < qual:                                       # :398
            lowq += 1
            continue
        if cell[r] == CELL_INVALID:                              # :412
            invalid_barcode += 1
            continue
        bc = cell[r]
        key = (bc, umi[r])
        if chrom[r] == CHROM_SC_SKIP:                            # :432  silent skip
            continue
        left = start[r]
        rite = end[r]
        if strand:
            cs = (chrom[r], S_MINUS if (flag[r] & F_REVERSE) else S_PLUS)   # :438
        else:
            cs = (chrom[r], S_NA)
        frag = (cs[0], cs[1], left, rite)
        if key in umis:                                          # :444
            first = next(iter(umis[key]))                        # :452 canonical: first inserted
            if (first[0], first[1]) == cs:
                already_seen += 1
                continue
            umis[key].setdefault(frag, None)                     # :459 set.add (string dedup)
            barcodes[bc] = barcodes.get(bc, 0) + 1               # :460-462
        else:
            umis[key] = {frag: None}                             # :470
            barcodes[bc] = barcodes.get(bc, 0) + 1               # :471-473
    if len(umis) > 0:                                            # :479
        save_bundle(umis)

    # ---- Part 2 (te_count.py:494-575)
    n_raw_barcodes = len(barcodes)
    todo = set(i[0] for i in sorted(barcodes.items(), key=itemgetter(1), reverse=True)[0:maxcells + pad])
    todo = sorted(todo, reverse=True)                            # :503
    handles = []
    for b in bundles:                                            # :509-513
        handles.append({"it": iter(b), "line": None, "BC": None, "open": True})
        handles[-1]["line"] = next(handles[-1]["it"])
        handles[-1]["BC"] = handles[-1]["line"][0]
    merged = []                                                  # lines of the merged .bun
    umi_count = 0
    while todo:
        cur = todo.pop()
        this_data = []
        for h in handles:
            while h["BC"] <= cur:                                # :528
                if not h["open"]:
                    break
                try:
                    h["line"] = next(h["it"])                    # :533 the held line is never kept
                    h["BC"] = h["line"][0]
                    if h["BC"] == cur:
                        this_data.append(h["line"])
                except StopIteration:
                    h["open"] = False
        umi_count += len(this_data)                              # :545
        seen = {}
        for line in this_data:
            k = (cur, line[1])
            if k not in seen:                                    # :552  first bundle wins,
                seen[k] = line[2]                                # :555  the `|` is discarded
        for k in seen:
            merged.append((k[0], k[1], seen[k]))

    # ---- Part 3 (te_count.py:577-707)
    triples = defaultdict(int)
    cell_hits = {}
    assigned = 0
    for (bc, _u, frags) in merged:
        reads = {}
        for f in frags:                                          # :603-606 later one wins
            reads[(f[0], f[1])] = (f[2], f[3])
        for (c, s), (left, rite) in reads.items():
            if c not in idx.buckets:                             # :614
                continue
            left_buck = ((left - 1) // bs) * bs                  # :619
            right_buck = (rite // bs) * bs                       # :620
            buckets_reqd = range(left_buck, right_buck + bs, bs)
            result = []
            loc_ids = set()
            if buckets_reqd:                                     # :625
                for buck in buckets_reqd:
                    if buck in idx.buckets[c]:
                        loc_ids.update(idx.buckets[c][buck])
                for i in loc_ids:
                    if left + 1 >= idx.L[i] and left <= idx.R[i]:        # :645
                        result.append(i)
                    if rite >= idx.L[i] and rite - 1 <= idx.R[i]:        # :648
                        result.append(i)
                if result:
                    cell_hits[bc] = cell_hits.get(bc, 0) + 1     # :653-655
                    types = set(idx.type_code[i] for i in result)
                    if any(idx.strand_code[i] == STRAND_MISSING for i in result):
                        raise ReferenceCrash("KeyError", "'strand' (te_count.py:661)")
                    ensgs = set((idx.ensg_id[i], idx.strand_code[i]) for i in result)   # :661
                    if T_GENE in types:
                        for e in ensgs:
                            if strand and s != e[1]:             # :665
                                continue
                            triples[(e[0], bc)] += 1
                    elif T_TE in types:
                        for e in ensgs:                          # :673 no strand filter
                            triples[(e[0], bc)] += 1
                    elif T_ENH in types:
                        for e in ensgs:                          # :679
                            triples[(e[0], bc)] += 1
                    else:
                        continue                                 # :684
                    assigned += 1
    if umi_count == 0:
        raise ReferenceCrash("ZeroDivisionError", "te_count.py:703")
    return {"triples": dict(triples), "cell_hits": list(cell_hits.items()),
            "stats": {"total_reads": units, "invalid_barcode": invalid_barcode,
                      "already_seen": already_seen, "lowq": lowq, "qcfail": qcfail,
                      "valid": umi_count, "assigned": assigned, "raw_barcodes": n_raw_barcodes,
                      "n_bundles": len(bundles)}}


def sc_select_cells(cell_hits, maxcells):
    """te_count.py:724-733 -- stable sort by hits descending, keep maxcells."""
    order = sorted(cell_hits, key=itemgetter(1), reverse=True)
    if len(cell_hits) > maxcells:
        order = order[0:maxcells]
    return [c for c, _ in order]
