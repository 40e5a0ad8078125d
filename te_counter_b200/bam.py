"""
Minimal BAM / SAM reader with the slice of pysam's interface that the counting path touches
(reference te_count/te_count.py:65-98, :190-214, :351-438): `AlignmentFile(filename, 'r')`, iteration,
`close()`, and per record `is_unmapped, is_duplicate, is_qcfail, is_reverse, mapping_quality,
query_name, reference_name, reference_start, reference_end, get_tags()`.

The reference (and `measureTE` here) uses pysam when it is installed; this module is the stand-in
for machines without it, so that a BAM file can be counted at all.  It follows the SAM/BAM
specification (SAMv1 section 4): BGZF is a series of gzip members, records are little-endian,
`reference_end` = `pos` + the reference-consuming CIGAR operations (M, D, N, =, X) and is None
for unmapped records or records without a CIGAR, as in pysam.  Parity of BAM decoding against
pysam itself is not pinned (pysam is absent here); tests/test_bam_reader.py checks
this reader against files written by an independent BAM writer.
"""
import gzip
import struct

_CIGAR_REF = (1, 0, 1, 1, 0, 0, 0, 1, 1)        # M I D N S H P = X : consumes reference?
_CIGAR_CHARS = "MIDNSHP=X"
_AUX_FMT = {b"c": "<b", b"C": "<B", b"s": "<h", b"S": "<H", b"i": "<i", b"I": "<I", b"f": "<f"}
_AUX_SIZE = {b"c": 1, b"C": 1, b"s": 2, b"S": 2, b"i": 4, b"I": 4, b"f": 4}


class AlignedSegment:
    __slots__ = ("flag", "mapping_quality", "query_name", "reference_name", "reference_start",
                 "reference_end", "_aux", "_aux_text")

    @property
    def is_unmapped(self):
        return bool(self.flag & 0x4)

    @property
    def is_reverse(self):
        return bool(self.flag & 0x10)

    @property
    def is_qcfail(self):
        return bool(self.flag & 0x200)

    @property
    def is_duplicate(self):
        return bool(self.flag & 0x400)

    def get_tags(self):
        """[(tag, value), ...] in file order"""
        if self._aux_text is not None:
            return _parse_sam_tags(self._aux_text)
        return _parse_bam_aux(self._aux)


def _parse_bam_aux(b):
    out = []
    i, n = 0, len(b)
    while i + 3 <= n:
        tag = b[i:i + 2].decode("ascii")
        t = b[i + 2:i + 3]
        i += 3
        if t in _AUX_FMT:
            out.append((tag, struct.unpack_from(_AUX_FMT[t], b, i)[0]))
            i += _AUX_SIZE[t]
        elif t == b"A":
            out.append((tag, b[i:i + 1].decode("ascii")))
            i += 1
        elif t in (b"Z", b"H"):
            j = b.index(b"\0", i)
            out.append((tag, b[i:j].decode("ascii")))
            i = j + 1
        elif t == b"B":
            st = b[i:i + 1]
            cnt = struct.unpack_from("<I", b, i + 1)[0]
            i += 5
            sz = _AUX_SIZE[st]
            out.append((tag, list(struct.unpack_from("<%d%s" % (cnt, _AUX_FMT[st][1]), b, i))))
            i += sz * cnt
        else:
            raise ValueError("unknown BAM aux type %r" % t)
    return out


def _parse_sam_tags(fields):
    out = []
    for f in fields:
        tag, t, v = f.split(":", 2)
        if t == "i":
            v = int(v)
        elif t == "f":
            v = float(v)
        elif t == "B":
            p = v.split(",")
            v = [float(x) if p[0] == "f" else int(x) for x in p[1:]]
        out.append((tag, v))
    return out


class AlignmentFile:
    def __init__(self, filename, mode="r"):
        self.filename = filename
        with open(filename, "rb") as fh:
            magic = fh.read(4)
        self.references = []
        if magic[:2] == b"\x1f\x8b":
            self._fh = gzip.open(filename, "rb")               # BGZF = concatenated gzip members
            if self._fh.read(4) != b"BAM\1":
                raise ValueError("%s: gzip-compressed but not a BAM file" % filename)
            self._read_bam_header()
            self._it = self._iter_bam()
        else:
            self._fh = open(filename, "r")
            self._it = self._iter_sam()

    # ------------------------------------------------------------------ BAM
    def _read_exact(self, n):
        b = self._fh.read(n)
        if len(b) != n:
            raise EOFError("truncated BAM file %s" % self.filename)
        return b

    def _read_bam_header(self):
        l_text = struct.unpack("<i", self._read_exact(4))[0]
        self.text = self._read_exact(l_text).rstrip(b"\0").decode("ascii", "replace")
        n_ref = struct.unpack("<i", self._read_exact(4))[0]
        for _ in range(n_ref):
            l_name = struct.unpack("<i", self._read_exact(4))[0]
            self.references.append(self._read_exact(l_name).rstrip(b"\0").decode("ascii"))
            self._read_exact(4)                                # l_ref

    def _iter_bam(self):
        fh, refs = self._fh, self.references
        unpack_head = struct.Struct("<iiBBHHHiiii").unpack_from
        while True:
            b = fh.read(4)
            if not b:
                return
            if len(b) != 4:
                raise EOFError("truncated BAM file %s" % self.filename)
            rec = self._read_exact(struct.unpack("<i", b)[0])
            ref_id, pos, l_name, mapq, _bin, n_cig, flag, l_seq, _nref, _npos, _tlen = unpack_head(rec, 0)
            r = AlignedSegment()
            r.flag = flag
            r.mapping_quality = mapq
            r.query_name = rec[32:32 + l_name - 1].decode("ascii")
            r.reference_name = refs[ref_id] if 0 <= ref_id < len(refs) else None
            r.reference_start = pos
            o = 32 + l_name
            end = None
            if n_cig and not (flag & 0x4):
                span = 0
                for op in struct.unpack_from("<%dI" % n_cig, rec, o):
                    if _CIGAR_REF[op & 0xF] if (op & 0xF) < 9 else 0:
                        span += op >> 4
                end = pos + (span or 1)          # htslib bam_endpos: pos + 1 when the CIGAR consumes no reference base
            r.reference_end = end
            o += 4 * n_cig + (l_seq + 1) // 2 + l_seq
            r._aux = rec[o:]
            r._aux_text = None
            yield r

    # ------------------------------------------------------------------ SAM text
    def _iter_sam(self):
        for line in self._fh:
            if line.startswith("@"):
                if line.startswith("@SQ"):
                    for f in line.rstrip("\n").split("\t")[1:]:
                        if f.startswith("SN:"):
                            self.references.append(f[3:])
                continue
            f = line.rstrip("\n").split("\t")
            if len(f) < 11:
                continue
            r = AlignedSegment()
            r.query_name = f[0]
            r.flag = int(f[1])
            r.reference_name = None if f[2] == "*" else f[2]
            r.reference_start = int(f[3]) - 1
            r.mapping_quality = int(f[4])
            end = None
            if f[5] != "*" and not (r.flag & 0x4):
                span, num = 0, 0
                for ch in f[5]:
                    if ch.isdigit():
                        num = num * 10 + ord(ch) - 48
                    else:
                        if ch in "MDN=X":
                            span += num
                        num = 0
                end = r.reference_start + (span or 1)
            r.reference_end = end
            r._aux = None
            r._aux_text = f[11:]
            yield r

    def __iter__(self):
        return self

    def __next__(self):
        return next(self._it)

    def close(self):
        self._fh.close()
