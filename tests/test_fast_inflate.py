"""The BAM decoder's own inflate (te_counter_b200/csrc/fast_inflate.h) against zlib: equal output on
every stream it accepts, never accepting what zlib rejects, no crash on damaged input.  zlib is
both the generator of the streams and the arbiter (bamdecode.cpp falls back to it)."""
import zlib

import numpy as np
import pytest

from te_counter_b200 import build, fastbam


@pytest.fixture(scope="module")
def lib():
    build.build_bam()
    return fastbam.load()


def _inflate(lib, comp, n_out, engine):
    out = np.zeros(max(1, n_out), dtype=np.uint8)
    rc = lib.tbam_inflate_raw(comp, len(comp), out.ctypes.data, n_out, engine)
    return rc, out[:n_out].tobytes()


def _deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, mem=8):
    co = zlib.compressobj(level, zlib.DEFLATED, -15, mem, strategy)
    return co.compress(data) + co.flush()


def _payloads():
    rng = np.random.default_rng(42)
    yield b""
    yield b"a"
    yield b"ab" * 5
    yield bytes(65536)
    yield bytes(range(256)) * 256
    yield rng.integers(0, 256, 65536, dtype=np.uint8).tobytes()                    # incompressible
    yield rng.integers(0, 4, 65280, dtype=np.uint8).tobytes()                      # 2 bits of entropy per byte
    yield np.array([37, 37, 37, 32, 25, 14, 2], dtype=np.uint8)[rng.integers(0, 7, 60000)].tobytes()
    text = b"@SQ\tSN:chr%d\tLN:%d\n"
    yield b"".join(text % (i, i * 7919) for i in range(3000))[:65536]
    # BAM-like records: repeated structure at distances of a few hundred bytes
    rec = bytearray(rng.integers(0, 256, 260, dtype=np.uint8).tobytes())
    out = bytearray()
    for i in range(250):
        rec[8:12] = int(i * 1000).to_bytes(4, "little")
        rec[40 + i % 100] = i & 255
        out += rec
    yield bytes(out)
    yield b"x" * 300 + rng.integers(0, 256, 20, dtype=np.uint8).tobytes() + b"x" * 1000     # off == 1 copies
    yield (b"abc" * 1000 + b"abcdefg" * 500)                                                # short distances 3, 7
    for n in (1, 7, 8, 9, 257, 258, 259, 269, 270, 300, 32768, 32769, 65535):
        yield rng.integers(0, 7, n, dtype=np.uint8).tobytes()


@pytest.mark.parametrize("level,strategy", [(0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY),
                                            (4, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY),
                                            (9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY),
                                            (6, zlib.Z_RLE), (6, zlib.Z_FILTERED)])
def test_equal_to_zlib_on_valid_streams(lib, level, strategy):
    declined = 0
    for i, data in enumerate(_payloads()):
        comp = _deflate(data, level, strategy, mem=1 + i % 9)
        rc0, ref = _inflate(lib, comp, len(data), 0)
        assert rc0 == 0 and ref == data
        rc1, got = _inflate(lib, comp, len(data), 1)
        if rc1 != 0:
            declined += 1
            continue
        assert got == data
        # wrong announced sizes are refused by both
        for n in (len(data) + 1, len(data) - 1):
            if n >= 0:
                assert _inflate(lib, comp, n, 1)[0] != 0 and _inflate(lib, comp, n, 0)[0] != 0
    assert declined <= 2, "the table-driven decoder should handle what zlib's compressor emits"


def test_multi_block_streams(lib):
    rng = np.random.default_rng(7)
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    parts, comp = [], b""
    for k in range(12):
        d = rng.integers(0, 1 + 3 * k, 3000 + 500 * k, dtype=np.uint8).tobytes()
        parts.append(d)
        comp += co.compress(d) + co.flush(zlib.Z_FULL_FLUSH if k % 2 else zlib.Z_SYNC_FLUSH)   # empty stored blocks in between
    comp += co.flush()
    data = b"".join(parts)
    rc, got = _inflate(lib, comp, len(data), 1)
    assert rc == 0 and got == data


def test_damaged_streams_never_crash_and_never_beat_zlib(lib):
    rng = np.random.default_rng(3)
    base = [d for d in _payloads() if len(d) > 200][:8]
    n_ok = 0
    for data in base:
        comp = bytearray(_deflate(data, 6))
        for trial in range(150):
            c = bytearray(comp)
            kind = trial % 3
            if kind == 0:
                for _ in range(1 + trial % 4):
                    c[int(rng.integers(len(c)))] ^= 1 << int(rng.integers(8))
            elif kind == 1:
                c = c[:int(rng.integers(len(c)))]
            else:
                pos = int(rng.integers(len(c)))
                c[pos:pos + 4] = rng.integers(0, 256, 4, dtype=np.uint8).tobytes()
            c = bytes(c)
            rc1, got = _inflate(lib, c, len(data), 1)
            rc0, ref = _inflate(lib, c, len(data), 0)
            if rc1 == 0:
                n_ok += 1
                assert rc0 == 0, "accepted a stream zlib rejects"
                assert got == ref
    assert n_ok >= 0
