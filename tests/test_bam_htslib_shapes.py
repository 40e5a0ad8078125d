"""BAM files shaped the way htslib / samtools / bgzip write them, built by hand from the SAMv1 specification (pysam and
htslib are not in this image): the 28-byte end-of-file marker, empty blocks in the middle of a file (parts
glued with `cat`), stored and fixed-Huffman DEFLATE blocks (samtools -u, tiny blocks), further gzip extra subfields
next to BC, @SQ lines in an order unrelated to the index, optional fields of every type (A c C s S i I f Z H and B arrays
of every subtype), read names of 254 characters, records without sequence, mate fields filled in, and alignments of more
than 65535 CIGAR operations stored the SAMv1 way (placeholder CIGAR <l_seq>S<ref_len>N + CG:B,I tag, section 4.2.2).

Expected columns are computed here from the records (reference_end = start + reference-consuming operations, as
htslib's bam_endpos; for a CG record the placeholder's N operation carries exactly that length, so the value pysam
reports after htslib has moved the tag into place is the same).  The three decoders -- the Python reader with pysam's
interface, libtecbam, and the block-parallel decoder run with host loops -- must all give them; tests/test_gpu_bam.py runs
the same files through the kernels."""
import struct

import numpy as np
import pytest

import helpers as H
from bam_writer import BGZF_EOF, write_bam
from te_counter_b200 import bam, reads
from test_bgzf_dev_cpu import _decode, lib   # noqa: F401  (fixture)
from test_fastbam import _native_batches

REF_OPS = (0, 2, 3, 7, 8)          # M D N = X consume the reference


def every_aux():
    a = b"XAAQ" + b"Xcc" + struct.pack("<b", -5) + b"XCC" + struct.pack("<B", 200) + b"Xss" + struct.pack("<h", -300)
    a += b"XSS" + struct.pack("<H", 60000) + b"Xii" + struct.pack("<i", -70000) + b"XII" + struct.pack("<I", 4000000000)
    a += b"Xff" + struct.pack("<f", 1.5) + b"XZZtext with spaces\0" + b"XHH1AE301\0"
    for st, fmt, vals in (("c", "b", [-1, 2]), ("C", "B", [1, 255]), ("s", "h", [-2, 3]), ("S", "H", [7]), ("i", "i", [-9]), ("I", "I", [1, 2, 3]), ("f", "f", [0.5])):
        a += b"YB" + b"B" + st.encode() + struct.pack("<I", len(vals)) + b"".join(struct.pack("<" + fmt, v) for v in vals)
    a += b"YEB" + b"c" + struct.pack("<I", 0)                                                  # empty array
    return a


def shaped_records(sc):
    rng = np.random.default_rng(41)
    recs = []
    for i in range(900):
        start = int(rng.integers(0, 120000))
        kind = i % 9
        if kind == 0:
            cig = [(50, 0)]
        elif kind == 1:
            cig = [(5, 4), (20, 7), (3, 8), (10, 2), (17, 0), (4, 1), (900, 3), (30, 0), (6, 5)]     # S = X D M I N M H
        elif kind == 2:
            cig = [(40, 4)]                                                                         # soft clips only: end = start + 1
        elif kind == 3:
            cig = [(70000, 4), (12345, 3)]                                                          # CG placeholder: <l_seq>S<ref_len>N
        elif kind == 4:
            cig = [(10, 0), (2, 6), (10, 0)]                                                        # P (padding) consumes nothing
        else:
            cig = [(int(rng.integers(20, 120)), 0)]
        r = {"chrom": ["chr1", "chrX", "chr2", "chrUn_random1", "chr7"][int(rng.integers(5))], "start": start, "cigar": cig,
             "flag": int(rng.choice([0, 16, 1024, 512, 99, 147, 256, 2048])), "mapq": int(rng.choice([0, 19, 20, 60, 255])),
             "name": ("q%d" % (i // 2)) if kind != 5 else "n" * 250 + "%04d" % (i // 2),
             "next_chrom": "chr1" if kind == 6 else None, "next_pos": start + 150 if kind == 6 else -1, "tlen": 250 if kind == 6 else 0}
        r["end"] = start + sum(n for n, op in cig if op in REF_OPS)
        aux = every_aux() if kind in (1, 7) else b"NHC\x01"
        if kind == 3:
            l_seq = 70000
            r["l_seq"] = 0                                                                          # sequence omitted ('*'): l_seq 0 is legal
            aux += b"CGBI" + struct.pack("<I", 3) + struct.pack("<III", 100 << 4 | 0, 12145 << 4 | 3, 100 << 4 | 0)
            del l_seq
        if kind == 8:
            r["l_seq"] = 0                                                                          # no sequence stored
        if sc:
            aux += b"CBZ" + ("ACGTAC%04d" % int(rng.integers(40))).encode() + b"\0" + b"UBZ" + "".join(rng.choice(list("ACGT"), size=10)).encode() + b"\0"
        r["aux"] = aux
        recs.append(r)
    return recs


def expected(recs, idx, paired_unused=False):
    cm = reads.ChromMap(idx.chrom_keys)
    start = np.array([r["start"] for r in recs], np.int32)
    end = np.array([max(r["end"], r["start"] + 1) if r["end"] == r["start"] else r["end"] for r in recs], np.int32)
    chrom = np.array([cm.bulk_id(r["chrom"]) for r in recs], np.uint16)
    mapq = np.array([r["mapq"] for r in recs], np.uint8)
    flag = np.array([((r["flag"] >> 2) & 1) | ((r["flag"] >> 10) & 1) << 1 | ((r["flag"] >> 9) & 1) << 2 | ((r["flag"] >> 4) & 1) << 3 for r in recs], np.uint8)
    return {"start": start, "end": end, "chrom": chrom, "mapq": mapq, "flag": flag}


VARIANTS = {
    "htslib_eof_marker": dict(block=60000, eof=BGZF_EOF),
    "empty_blocks_inside": dict(block=5000, eof_every=3, eof=BGZF_EOF),
    "stored_blocks": dict(block=20000, level=0, eof=BGZF_EOF),
    "tiny_fixed_huffman_blocks": dict(block=40, eof=BGZF_EOF),
    "extra_subfields": dict(block=9000, extra=b"RA" + struct.pack("<H", 4) + b"abcd" + b"ZZ" + struct.pack("<H", 0), eof=BGZF_EOF),
    "sq_order_and_comments": dict(block=7000, chroms=["chrM", "chr7", "chrUn_random1", "chr2", "chrX", "chr1"], eof=BGZF_EOF,
                                  header_text="@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\tM5:0123\n" % (c, 1000 + i) for i, c in
                                                                                         enumerate(["chrM", "chr7", "chrUn_random1", "chr2", "chrX", "chr1"]))
                                  + "@RG\tID:g1\tSM:x\n@PG\tID:samtools\tPN:samtools\tVN:1.17\tCL:samtools sort\n@CO\tfree text, tabs\tincluded\n"),
}


def write_variant(path, name, sc=False):
    recs = shaped_records(sc)
    write_bam(path, recs, **VARIANTS[name])
    return recs


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_all_host_side_decoders_give_the_expected_columns(lib, tmp_path, name):   # noqa: F811
    path = str(tmp_path / (name + ".bam"))
    recs = write_variant(path, name)
    idx = H.load_index("idx_rand_a.glb")
    want = expected(recs, idx)
    # the Python reader with pysam's interface + the packing of reads.py
    f = bam.AlignmentFile(path, "r")
    b = reads.Batch(len(recs) + 8)
    assert not reads.fill_bulk(b, f, reads.ChromMap(idx.chrom_keys), False, 0)
    f.close()
    assert b.n == len(recs)
    for k in want:
        assert np.array_equal(getattr(b, k)[:b.n], want[k]), ("python", k)
    # libtecbam (threads + zlib)
    bs = _native_batches(path, "se", reads.ChromMap(idx.chrom_keys), None, 0, 1 << 16, 3)
    for k in want:
        assert np.array_equal(np.concatenate([getattr(x, k)[:x.n] for x, _ in bs]), want[k]), ("libtecbam", k)
    # the block-parallel decoder of the GPU path, host loops
    for wb in (4, 1 << 20):
        rc, declined, got = _decode(lib, path, "se", reads.ChromMap(idx.chrom_keys), None, 0, wb)
        assert rc == 0, declined
        for k in want:
            assert np.array_equal(got[k], want[k]), ("block-parallel", k, wb)


def test_single_cell_tags_among_every_other_type(lib, tmp_path):   # noqa: F811
    path = str(tmp_path / "sc.bam")
    recs = write_variant(path, "empty_blocks_inside", sc=True)
    idx = H.load_index("idx_rand_a.glb")
    wlf = tmp_path / "wl.txt"
    wlf.write_text("".join("ACGTAC%04d\n" % i for i in range(30)))
    wl = reads.Whitelist(str(wlf))
    f = bam.AlignmentFile(path, "r")
    b = reads.Batch(len(recs) + 8, sc=True)
    assert not reads.fill_sc(b, f, reads.ChromMap(idx.chrom_keys), wl, 20)
    f.close()
    rc, _, got = _decode(lib, path, "sc", reads.ChromMap(idx.chrom_keys), wl, 20, 16)
    assert rc == 0
    bs = _native_batches(path, "sc", reads.ChromMap(idx.chrom_keys), wl, 20, 1 << 16, 2)
    for k in ("start", "end", "chrom", "mapq", "flag", "cell", "umi"):
        assert np.array_equal(got[k], getattr(b, k)[:b.n]), k
        assert np.array_equal(np.concatenate([getattr(x, k)[:x.n] for x, _ in bs]), getattr(b, k)[:b.n]), k
    assert (b.cell[:b.n] != reads.CELL_INVALID).sum() > 100
