"""TEST INFRASTRUCTURE: corrupted record streams (valid BGZF container, damaged BAM records) through the
block-parallel decoder built with AddressSanitizer (host-loop backend, tools/bgzf_dev_host.cpp): no read or
write outside the buffers, only status codes.  Run:

    g++ -O1 -g -std=c++17 -fsanitize=address -fno-omit-frame-pointer -shared -fPIC tools/bgzf_dev_host.cpp -lz \
        -o /tmp/libbgzfdevhost_asan.so
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tools/bam_corrupt_fuzz.py

Round 1: 300 files x 2 modes x 3 window sizes = 1800 runs, no sanitizer report; status histogram
{-2 (format): 930, -113 (reference_end None on a counted record): 432, 0: 356, -102 (malformed record): 39,
 -5 (split refused): 43}.
"""
import sys, os, ctypes, tempfile, struct
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from bam_writer import _bgzf_block
so = ctypes.CDLL('/tmp/libbgzfdevhost_asan.so'); so.bgzfdev_error.restype = ctypes.c_char_p
import test_bam_fuzz_cpu as T
tmp = tempfile.mkdtemp()
stats = {}
for seed in range(300):
    rng = np.random.default_rng(1000 + seed)
    n_ref = int(rng.choice([1, 25, 3000]))
    path = os.path.join(tmp, "f.bam")
    T._random_bam(path, rng, 200, n_ref)
    # inflate all, mutate the uncompressed bytes behind the header, re-block
    import zlib
    raw = open(path, 'rb').read(); o = 0; data = b""
    while o < len(raw):
        n = int.from_bytes(raw[o+16:o+18], 'little') + 1
        data += zlib.decompress(raw[o+18:o+n-8], -15); o += n
    data = bytearray(data)
    hdr_end = 12 + struct.unpack_from("<i", data, 4)[0]
    nref = struct.unpack_from("<i", data, hdr_end - 4)[0]
    p = hdr_end
    for _ in range(nref):
        p += 8 + struct.unpack_from("<i", data, p)[0]
    for _ in range(int(rng.integers(1, 6))):
        k = int(rng.integers(p, len(data)))
        kind = int(rng.integers(3))
        if kind == 0: data[k] ^= 1 << int(rng.integers(8))
        elif kind == 1: data[k:k+4] = rng.integers(0, 256, 4, dtype=np.uint8).tobytes()
        else: del data[k:k+int(rng.integers(1, 50))]
    with open(path, 'wb') as fh:
        o = 0
        while o < len(data):
            n = int(rng.choice([60, 300, 3000, 65000])); fh.write(_bgzf_block(bytes(data[o:o+n]))); o += n
        fh.write(_bgzf_block(b""))
    for mode in (0, 1):
        for wb in (1, 4, 1 << 16):
            h = ctypes.c_void_p()
            if so.bgzfdev_open(path.encode(), ctypes.byref(h)) != 0: continue
            nr = so.bgzfdev_n_references(h)
            ids = np.arange(nr, dtype=np.uint16)
            so.bgzfdev_set_chrom_map(h, ids.ctypes.data_as(ctypes.c_void_p), ids.ctypes.data_as(ctypes.c_void_p), nr, nr)
            n = ctypes.c_int64(0)
            rc = so.bgzfdev_decode(h, mode, 0, wb, ctypes.byref(n))
            stats[rc] = stats.get(rc, 0) + 1
            so.bgzfdev_close(h)
print("status histogram over corrupted files:", stats)
