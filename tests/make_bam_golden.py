"""Writes the committed BAM decoding fixtures under tests/golden/: small BGZF files with every
record kind the packing distinguishes (tests/test_fastbam._mixed_records) and, beside them, the
arrays the Python packing (te_counter_b200/bam.py + reads.py, the restatement of the reference's
read loop) produces from them.  The decoders are tested against these vectors.

    python tests/make_bam_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import helpers as H                                     # noqa: E402
from bam_writer import write_bam                        # noqa: E402
from te_counter_b200 import bam, reads                 # noqa: E402
from test_fastbam import _mixed_records                 # noqa: E402


def main():
    idx = H.load_index("idx_rand_a.glb")
    out = {}
    for mode, seed, block in (("se", 101, 700), ("pe", 102, 3000), ("sc", 103, 1500)):
        recs, wl_list = _mixed_records(601 if mode == "pe" else 600, seed, mode == "sc")
        path = os.path.join(H.GOLD, "bam_mixed_%s.bam" % mode)
        write_bam(path, recs, block=block)
        wl = None
        if mode == "sc":
            wlp = os.path.join(H.GOLD, "bam_mixed_whitelist.txt")
            with open(wlp, "w") as fh:
                fh.write("".join(w + "\n" for w in wl_list))
            wl = reads.Whitelist(wlp)
        f = bam.AlignmentFile(path, "r")
        b = reads.Batch(1024, sc=mode == "sc")
        cm = reads.ChromMap(idx.chrom_keys)
        more = reads.fill_sc(b, f, cm, wl, 20) if mode == "sc" else reads.fill_bulk(b, f, cm, mode == "pe", 20)
        assert not more and b.n == 600
        f.close()
        for k in ("start", "end", "chrom", "mapq", "flag") + (("cell", "umi") if mode == "sc" else ()):
            out["%s_%s" % (mode, k)] = getattr(b, k)[:b.n].copy()
    np.savez_compressed(os.path.join(H.GOLD, "bam_mixed_expected.npz"), **out)
    print("written:", sorted(out))


if __name__ == "__main__":
    main()
