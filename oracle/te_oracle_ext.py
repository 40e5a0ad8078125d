"""
TEST INFRASTRUCTURE -- CPU restatement of the OPT-IN EXTENSIONS of te_counter_b200.

These are defined semantics for flag combinations on which the reference raises; they are NOT part of the
parity claim (there is nothing in the reference to be equal to).  Each function states the rule it implements and
the reference site that raises instead.  Same import rules as oracle/te_oracle.py: tests only.

  bulk --strand      te_count/te_count.py:58-59 and :183-184 raise NotImplementedError.
                     Rule: the unit's strand is that of its first record (flag 0x10 set = '-'); a feature whose
                     strand is '+' or '-' is a candidate only for units on that strand, a feature with any other
                     strand value (or none) for both.  Filters, the two-bucket candidate rule, the point tests, the
                     type rule, the tally and the statistics are those of the unstranded loop.
  -q N               bin/te_count:30 declares -q with nargs=1, so measureTE receives [N] and te_count.py:88
                     raises TypeError.  Rule: a one-element list or tuple means its element.
  --noumi            te_count.py:429-442 records nothing when UMIS is False and :703 divides by zero.
                     Rule: every surviving record is its own molecule -- its UMI is its ordinal in the file -- and
                     the rest of the pipeline (bundles, top cells, held-line drop, overlap, tally) is unchanged.
"""
from . import te_oracle
from .te_oracle import F_REVERSE


def defined_quality(q):
    """-q N: [N] -> N."""
    if isinstance(q, (list, tuple)) and len(q) == 1:
        return int(q[0])
    return int(q)


def bulk_count_stranded(idx, paired, qual, start, end, chrom, mapq, flag):
    """bulk --strand (see the module header)."""
    def candidate(i, r1):
        fs = idx.strand_code[i]
        us = 1 if (flag[r1] & F_REVERSE) else 0
        return not (fs in (0, 1) and fs != us)
    return te_oracle.bulk_count(idx, paired, qual, start, end, chrom, mapq, flag, candidate=candidate)


def sc_count_noumi(idx, qual, strand, bundle_keys, maxcells, pad, start, end, chrom, mapq, flag, cell):
    """--noumi (see the module header): the UMI of record i is i."""
    umi = list(range(len(start)))
    return te_oracle.sc_count(idx, qual, strand, bundle_keys, maxcells, pad, start, end, chrom, mapq, flag, cell, umi)
