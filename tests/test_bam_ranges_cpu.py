"""One BAM split into byte ranges (csrc/bam_orch.h decode_range; include/tecount.h tec_bam_count_range), run with host
loops in place of the kernels (tools/bgzf_dev_host.cpp): for any number of ranks the ranges' records, concatenated in
rank order, are the records of the whole file, every rank's exit equals the next rank's start, and the last exit is the
end of the file -- the check the multi-GPU callers make (te_counter_b200/shard.py)."""
import ctypes
import os

import numpy as np
import pytest

import helpers as H
from bam_writer import write_bam
from te_counter_b200 import reads, shard
from test_bgzf_dev_cpu import MODES, _decode, lib   # noqa: F401  (fixture)
from test_fastbam import _mixed_records


def _decode_range(so, path, mode, cm, wl, qual, window_blocks, lo, hi):
    h = ctypes.c_void_p()
    assert so.bgzfdev_open(path.encode(), ctypes.byref(h)) == 0
    try:
        refs = [so.bgzfdev_reference_name(h, i).decode() for i in range(so.bgzfdev_n_references(h))]
        bulk = np.array([cm.bulk_id(n) for n in refs], dtype=np.uint16)
        sc = np.array([cm.sc_id(n) for n in refs], dtype=np.uint16)
        assert so.bgzfdev_set_chrom_map(h, bulk.ctypes.data_as(ctypes.c_void_p), sc.ctypes.data_as(ctypes.c_void_p), len(refs), cm.n_index) == 0
        if wl is not None:
            enc = [b.encode() for b in wl.id_to_barcode]
            off = np.zeros(len(enc) + 1, dtype=np.int64)
            np.cumsum([len(b) for b in enc], out=off[1:])
            assert so.bgzfdev_set_whitelist(h, b"".join(enc), off.ctypes.data_as(ctypes.c_void_p), len(enc)) == 0
        out = (ctypes.c_int64 * 6)()
        rc = so.bgzfdev_decode_range(h, MODES[mode], qual, window_blocks, ctypes.c_int64(lo), ctypes.c_int64(hi), out)
        assert rc == 0, (rc, so.bgzfdev_error(h).decode())
        n = out[0]
        cols = {"start": np.zeros(n, np.int32), "end": np.zeros(n, np.int32), "chrom": np.zeros(n, np.uint16),
                "mapq": np.zeros(n, np.uint8), "flag": np.zeros(n, np.uint8), "cell": np.zeros(n, np.uint32), "umi": np.zeros(n, np.uint64)}
        so.bgzfdev_fetch(h, *[cols[k].ctypes.data_as(ctypes.c_void_p) for k in ("start", "end", "chrom", "mapq", "flag", "cell", "umi")])
        return {"n": n, "start": (out[1], out[2]), "exit": (out[3], out[4]), "size": out[5]}, cols
    finally:
        so.bgzfdev_close(h)


@pytest.mark.parametrize("mode", ["se", "sc"])
@pytest.mark.parametrize("block,window_blocks", [(3000, 1 << 20), (700, 5), (65000, 2), (300, 3)])
def test_ranges_concatenate_to_the_whole_file(lib, tmp_path, mode, block, window_blocks):   # noqa: F811
    recs, wl_list = _mixed_records(4000, 5 + block, mode == "sc")
    path = str(tmp_path / "x.bam")
    write_bam(path, recs, block=block)
    idx = H.load_index("idx_rand_a.glb")
    wl = None
    if mode == "sc":
        wlf = tmp_path / "wl.txt"
        wlf.write_text("".join(w + "\n" for w in wl_list))
        wl = reads.Whitelist(str(wlf))
    rc, _, whole = _decode(lib, path, mode, reads.ChromMap(idx.chrom_keys), wl, 20, window_blocks)
    assert rc == 0
    size = os.path.getsize(path)
    keys = ("start", "end", "chrom", "mapq", "flag") + (("cell", "umi") if mode == "sc" else ())
    for world in (1, 2, 3, 8, 37):
        infos, parts = [], []
        for rank in range(world):
            lo, hi = shard.byte_range(size, rank, world)
            info, cols = _decode_range(lib, path, mode, reads.ChromMap(idx.chrom_keys), wl, 20, window_blocks, lo, hi)
            infos.append(info)
            parts.append(cols)
        assert shard.chain_is_consistent(infos), (world, infos)
        for k in keys:
            assert np.array_equal(np.concatenate([p[k] for p in parts]), whole[k]), (k, world)
        if world > 1 and block <= 3000:
            assert sum(1 for i in infos if i["n"]) > 1                # the split really happened


def test_long_records_across_range_borders(lib, tmp_path):   # noqa: F811
    """Records longer than a block: most blocks hold no record start, ranks adopt a later block, tails reach over several blocks."""
    recs = [{"chrom": "chr1", "start": 100 + i, "end": 150 + i, "name": "n" * 200 + str(i)} for i in range(400)]
    path = str(tmp_path / "x.bam")
    write_bam(path, recs, block=97)
    cm = lambda: reads.ChromMap(["1"])   # noqa: E731
    rc, _, whole = _decode(lib, path, "se", cm(), None, 0, 64)
    assert rc == 0 and len(whole["start"]) == 400
    size = os.path.getsize(path)
    for world in (2, 5, 16):
        infos, parts = [], []
        for rank in range(world):
            lo, hi = shard.byte_range(size, rank, world)
            info, cols = _decode_range(lib, path, "se", cm(), None, 0, 7, lo, hi)
            infos.append(info)
            parts.append(cols)
        assert shard.chain_is_consistent(infos)
        assert np.array_equal(np.concatenate([p["start"] for p in parts]), whole["start"])


def test_chain_check_rejects_a_wrong_start():
    ok = [{"n": 5, "start": (-1, 0), "exit": (900, 12), "size": 2000}, {"n": 0, "start": (-2, 0), "exit": (-2, 0), "size": 2000},
          {"n": 7, "start": (900, 12), "exit": (2000, 0), "size": 2000}]
    assert shard.chain_is_consistent(ok)
    bad = [dict(ok[0]), dict(ok[1]), dict(ok[2], start=(900, 40))]
    assert not shard.chain_is_consistent(bad)
    assert not shard.chain_is_consistent([dict(ok[0]), dict(ok[2], exit=(1990, 3))])       # does not reach the end of the file
    assert not shard.chain_is_consistent([dict(ok[0], start=(10, 0)), ok[2]])              # rank 0 must start at the header
