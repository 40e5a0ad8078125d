"""TEST INFRASTRUCTURE: an independent BAM / SAM writer (SAMv1 section 4) used to check
te_counter_b200/bam.py and to run the golden cases from real files.  Records are the dicts of
tests/golden (chrom, start, end, mapq, flag, name, CB/CR/UB/UR); the alignment is written as
`<a>M <gap>N <b>M` so that reference_end has to be recovered from the CIGAR."""
import struct
import zlib


# the 28-byte end-of-file marker of SAMv1 section 4.1.2, as htslib / bgzip write it (an empty block, fixed Huffman code)
BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _bgzf_block(data, level=6, extra=b""):
    """One BGZF block.  level 0 gives stored DEFLATE blocks (samtools -u / bgzip -l 0); `extra` = further gzip extra
    subfields placed in front of BC (the format allows them; readers must skip what they do not know)."""
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    xlen = 6 + len(extra)
    bsize = len(comp) + 19 + xlen
    head = struct.pack("<BBBBIBBH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, xlen) + extra + struct.pack("<BBHH", ord("B"), ord("C"), 2, bsize)
    return head + comp + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data))


def _cigar_for(rec):
    if rec.get("cigar") is not None:                     # explicit list of (length, op) pairs
        return list(rec["cigar"])
    span = rec["end"] - rec["start"]
    if rec.get("flag", 0) & 0x4 or span <= 0:
        return []
    if span >= 30:
        a = span // 3
        gap = span // 3
        return [(a, 0), (gap, 3), (5, 1), (span - a - gap, 0), (7, 4)]      # M N I M S
    return [(span, 0)]


def _aux(rec):
    if rec.get("aux") is not None:                       # raw optional fields, as given
        return rec["aux"]
    out = b""
    for t in ("CB", "CR", "UB", "UR"):
        if rec.get(t) is not None:
            out += t.encode() + b"Z" + rec[t].encode() + b"\0"
    out += b"NHC" + struct.pack("<B", 1) + b"ASi" + struct.pack("<i", -3) + b"XBBs" + struct.pack("<Ihh", 2, -1, 5)
    return out


def write_bam(path, records, block=3000, level=6, extra=b"", eof_every=0, chroms=None, header_text=None, eof=None):
    """eof_every: an empty block after every n-th block (files glued together with `cat` carry them); chroms: the @SQ
    order (default: first appearance); eof: the bytes of the final marker (default: an empty block from zlib)."""
    chroms = list(chroms) if chroms is not None else []
    for r in records:
        if r["chrom"] is not None and r["chrom"] not in chroms:
            chroms.append(r["chrom"])
    text = header_text if header_text is not None else \
        "@HD\tVN:1.6\tSO:unsorted\n" + "".join("@SQ\tSN:%s\tLN:300000000\n" % c for c in chroms)
    raw = b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(chroms))
    for c in chroms:
        raw += struct.pack("<i", len(c) + 1) + c.encode() + b"\0" + struct.pack("<i", 300000000)
    for i, r in enumerate(records):
        name = r.get("name", "r%d" % i).encode() + b"\0"
        cig = _cigar_for(r)
        l_seq = r["l_seq"] if r.get("l_seq") is not None else sum(n for n, op in cig if op in (0, 1, 4, 7, 8))
        ref_id = chroms.index(r["chrom"]) if r["chrom"] is not None else -1
        next_id = chroms.index(r["next_chrom"]) if r.get("next_chrom") is not None else -1
        body = struct.pack("<iiBBHHHiiii", ref_id, r["start"], len(name), r.get("mapq", 60), 4680, len(cig),
                           r.get("flag", 0), l_seq, next_id, r.get("next_pos", -1), r.get("tlen", 0))
        body += name + b"".join(struct.pack("<I", n << 4 | op) for n, op in cig)
        body += b"\x11" * ((l_seq + 1) // 2) + b"\xff" * l_seq + _aux(r)
        raw += struct.pack("<i", len(body)) + body
    with open(path, "wb") as fh:
        for k, o in enumerate(range(0, len(raw), block)):    # many small blocks: records straddle them
            fh.write(_bgzf_block(raw[o:o + block], level, extra))
            if eof_every and (k + 1) % eof_every == 0:
                fh.write(BGZF_EOF)
        fh.write(_bgzf_block(b"") if eof is None else eof)   # EOF marker


def write_sam(path, records):
    chroms = []
    for r in records:
        if r["chrom"] is not None and r["chrom"] not in chroms:
            chroms.append(r["chrom"])
    with open(path, "w") as fh:
        fh.write("@HD\tVN:1.6\n" + "".join("@SQ\tSN:%s\tLN:300000000\n" % c for c in chroms))
        for i, r in enumerate(records):
            cig = "".join("%d%s" % (n, "MIDNSHP=X"[op]) for n, op in _cigar_for(r)) or "*"
            tags = ["%s:Z:%s" % (t, r[t]) for t in ("CB", "CR", "UB", "UR") if r.get(t) is not None] + ["NH:i:1"]
            fh.write("\t".join([r.get("name", "r%d" % i), str(r.get("flag", 0)), r["chrom"] or "*", str(r["start"] + 1),
                                str(r.get("mapq", 60)), cig, "*", "0", "0", "*", "*"] + tags) + "\n")
