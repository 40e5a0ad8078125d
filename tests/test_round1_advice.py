"""Round-1 advisor findings, each with the case that shows it (CPU side; the device decoder shares
bgzf_dev.h's record parser, which tests/test_bgzf_dev_cpu.py runs on the host)."""
import ctypes
import os

import numpy as np
import pytest

import helpers as H
from bam_writer import write_bam
from te_counter_b200 import _lib, bam, fastbam, reads


def test_bulk_chrom_map_takes_any_number_of_reference_names():
    """A header with more sequences than there are u16 ids (scaffold-level assemblies): bulk mode only asks
    "is it an index chromosome", so names outside the index share CHROM_INVALID and never run out of ids."""
    cm = reads.ChromMap(["1", "2", "X"])
    for i in range(70000):
        assert cm.bulk_id("scaffold_%d" % i) == reads.CHROM_INVALID
    assert cm.bulk_id("chr2") == 1 and cm.bulk_id("X") == 2 and cm.bulk_id(None) == reads.CHROM_INVALID
    # single cell still tells contigs apart (the chrom:strand comparison of te_count.py:446-452)
    assert cm.sc_id("scaffoldA") != cm.sc_id("scaffoldB")


@pytest.mark.parametrize("decoder", ["python", "native"])
def test_reference_end_of_a_cigar_without_reference_bases(tmp_path, decoder):
    """htslib's bam_endpos (pysam reference_end) is pos + 1 when no CIGAR operation consumes the reference
    (soft clips / insertions only); a record without CIGAR has no reference_end at all."""
    recs = [{"chrom": "chr1", "start": 1000, "end": 1001, "mapq": 60, "flag": 0, "cigar": [(20, 4), (5, 1)]},     # 20S5I
            {"chrom": "chr1", "start": 2000, "end": 2050, "mapq": 60, "flag": 0, "cigar": [(3, 4), (50, 0)]},    # 3S50M
            {"chrom": "chr1", "start": 3000, "end": 3010, "mapq": 60, "flag": 0, "cigar": [(4, 7), (6, 8)]}]     # 4=6X
    path = str(tmp_path / "x.bam")
    write_bam(path, recs)
    want = [(1000, 1001), (2000, 2050), (3000, 3010)]
    if decoder == "python":
        f = bam.AlignmentFile(path, "r")
        got = [(r.reference_start, r.reference_end) for r in f]
        f.close()
    else:
        f = fastbam.NativeBam(path, threads=1)
        f.bind(reads.ChromMap(["1"]), None)
        b = reads.Batch(16)
        f.fill_bulk(b, False, 0)
        got = list(zip(b.start[:b.n].tolist(), b.end[:b.n].tolist()))
        f.close()
    assert got == want


def test_strerror_covers_every_status():
    lib = _lib.load_library()
    lib.tec_strerror.restype = ctypes.c_char_p
    texts = {s: lib.tec_strerror(s).decode() for s in range(0, -16, -1)}
    assert all(t and t != "unknown status" for t in texts.values()), texts
    assert lib.tec_strerror(-99).decode() == "unknown status"
