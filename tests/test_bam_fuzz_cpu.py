"""Randomised BAM files with arbitrary (valid) record shapes -- names, CIGARs, sequence lengths,
optional fields of every type, many reference sequences -- cut into BGZF blocks at random sizes:
the host decoder (libtecbam), the block-parallel decoder (host-loop build of the device code) and
the Python reader must agree, for any window size, and the block-parallel split must not refuse."""
import struct

import numpy as np
import pytest

from bam_writer import _bgzf_block
from te_counter_b200 import bam, reads
from test_bgzf_dev_cpu import _decode, _native, lib       # noqa: F401  (lib is a fixture)
from test_fastbam import _canon_chrom


def _random_bam(path, rng, n_rec, n_ref):
    refs = ["ref%d%s" % (i, "_alt" if i % 11 == 0 else "") for i in range(n_ref)]
    text = "@HD\tVN:1.6\n" + "".join("@SQ\tSN:%s\tLN:1000000\n" % r for r in refs[:50])
    raw = bytearray(b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", n_ref))
    for r in refs:
        raw += struct.pack("<i", len(r) + 1) + r.encode() + b"\0" + struct.pack("<i", 1000000)
    tag_types = "AcCsSiIfZHB"
    for i in range(n_rec):
        name = bytes(rng.integers(0x21, 0x7F, int(rng.integers(1, 60)), dtype=np.uint8).tolist()).replace(b"@", b"a") + b"\0"
        n_cig = int(rng.choice([0, 1, 1, 2, 5, 40]))
        cig = [(int(rng.integers(1, 300)), int(rng.integers(0, 9))) for _ in range(n_cig)]
        l_seq = int(rng.choice([0, 1, 36, 100, 151, 1000]))
        flag = int(rng.integers(0, 1 << 12))
        ref_id = int(rng.integers(-1, n_ref))
        pos = int(rng.integers(-1, 900000))
        if ref_id < 0:
            flag |= 4
        aux = b""
        for _ in range(int(rng.integers(0, 6))):
            t = tag_types[int(rng.integers(len(tag_types)))]
            tag = bytes(rng.integers(65, 91, 2, dtype=np.uint8).tolist())
            if tag in (b"CB", b"CR", b"UB", b"UR"):
                tag = b"XX"
            aux += tag + t.encode()
            if t == "A":
                aux += b"q"
            elif t in "cC":
                aux += struct.pack("<B", int(rng.integers(0, 128)))
            elif t in "sS":
                aux += struct.pack("<H", int(rng.integers(0, 30000)))
            elif t in "iI":
                aux += struct.pack("<I", int(rng.integers(0, 1 << 31)))
            elif t == "f":
                aux += struct.pack("<f", float(rng.random()))
            elif t in "ZH":
                aux += bytes(rng.integers(0x30, 0x5B, int(rng.integers(0, 40)), dtype=np.uint8).tolist()) + b"\0"
            else:
                st = "cCsSiIf"[int(rng.integers(7))]
                cnt = int(rng.integers(0, 12))
                aux += st.encode() + struct.pack("<I", cnt) + bytes(cnt * {"c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}[st])
        body = struct.pack("<iiBBHHHiiii", ref_id, pos, len(name), int(rng.integers(0, 256)), 4680, n_cig, flag, l_seq,
                           int(rng.integers(-1, n_ref)), int(rng.integers(-1, 900000)), int(rng.integers(-500, 500)))
        body += name + b"".join(struct.pack("<I", n << 4 | op) for n, op in cig)
        body += bytes(rng.integers(0, 256, (l_seq + 1) // 2 + l_seq, dtype=np.uint8).tolist()) + aux
        raw += struct.pack("<i", len(body)) + body
    raw = bytes(raw)
    with open(path, "wb") as fh:
        o = 0
        while o < len(raw):
            n = int(rng.choice([60, 300, 3000, 20000, 65000]))
            fh.write(_bgzf_block(raw[o:o + n]))
            o += n
        fh.write(_bgzf_block(b""))


@pytest.mark.parametrize("seed,n_ref", [(1, 1), (2, 25), (3, 3000), (4, 40000), (5, 7), (6, 300)])
def test_random_record_shapes(lib, tmp_path, seed, n_ref):        # noqa: F811
    rng = np.random.default_rng(seed)
    path = str(tmp_path / "f.bam")
    _random_bam(path, rng, 1500, n_ref)
    keys = ["ref%d" % i for i in range(0, n_ref, 2)]
    # Python reader: reference_end and flags
    f = bam.AlignmentFile(path, "r")
    py = [(r.reference_start, -1 if r.reference_end is None else r.reference_end, r.mapping_quality,
           int(r.is_unmapped) | int(r.is_duplicate) << 1 | int(r.is_qcfail) << 2 | int(r.is_reverse) << 3) for r in f]
    f.close()
    assert len(py) == 1500
    # qual 256 > any MAPQ: nothing is "counted", so a record without reference_end never raises
    want = _native(path, "se", reads.ChromMap(keys), None, 256)
    assert want["start"].tolist() == [p[0] for p in py] and want["end"].tolist() == [p[1] for p in py]
    assert want["mapq"].tolist() == [p[2] for p in py] and want["flag"].tolist() == [p[3] for p in py]

    class _B:
        pass
    for wb in (1, 3, 17, 1 << 16):
        rc, msg, got = _decode(lib, path, "se", reads.ChromMap(keys), None, 256, wb)
        assert rc == 0, (wb, msg)
        a, b = _B(), _B()
        a.n = b.n = 1500
        a.chrom, b.chrom = want["chrom"], got["chrom"]
        assert np.array_equal(_canon_chrom([(a, False)], len(keys))[0], _canon_chrom([(b, False)], len(keys))[0])
        for k in ("start", "end", "mapq", "flag"):
            assert np.array_equal(want[k], got[k]), (k, wb)


@pytest.mark.parametrize("seed,n_ref,n_rec", [(11, 3, 1501), (12, 300, 1500), (13, 3000, 777)])
def test_random_record_shapes_paired(lib, tmp_path, seed, n_ref, n_rec):      # noqa: F811
    """Pairs in file order (mate-name rule, odd trailing record) across block and window borders."""
    rng = np.random.default_rng(seed)
    path = str(tmp_path / "f.bam")
    _random_bam(path, rng, n_rec, n_ref)
    keys = ["ref%d" % i for i in range(0, n_ref, 2)]
    want = _native(path, "pe", reads.ChromMap(keys), None, 0)
    assert len(want["start"]) == n_rec - (n_rec & 1)
    for wb in (1, 2, 9, 1 << 16):
        rc, msg, got = _decode(lib, path, "pe", reads.ChromMap(keys), None, 0, wb)
        assert rc == 0, (wb, msg)
        for k in ("start", "end", "mapq", "flag"):
            assert np.array_equal(want[k], got[k]), (k, wb)
