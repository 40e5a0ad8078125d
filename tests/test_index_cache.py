"""Sidecar of the flattened index (te_counter_b200/index.py): derived from the .glb,
used only while it matches the .glb on disk."""
import os
import shutil

import numpy as np
import pytest

import helpers as H
from te_counter_b200 import index as tindex

ARRAYS = ("chrom_id", "L", "R", "ensg_id", "type_code", "strand_code")


def _same(a, b):
    assert a.names == b.names and a.chrom_keys == b.chrom_keys and a.strand_strings == b.strand_strings
    assert a.bucket_size == b.bucket_size
    for k in ARRAYS:
        x, y = getattr(a, k), getattr(b, k)
        assert x.dtype == y.dtype and np.array_equal(x, y), k


@pytest.mark.parametrize("glb", ["idx_rand_a.glb", "idx_toy.glb", "idx_rand_sc.glb"])
def test_sidecar_round_trip_and_invalidation(monkeypatch, tmp_path, glb):
    path = str(tmp_path / glb)
    shutil.copy(os.path.join(H.GOLD, glb), path)
    side = path + tindex.CACHE_SUFFIX
    ref = tindex.load_glb(path, cache=False)
    assert not os.path.exists(side)
    first = tindex.load_glb(path, cache=True)
    assert os.path.exists(side)
    _same(ref, first)

    calls = []
    real_read = tindex._read_pickle
    monkeypatch.setattr(tindex, "_read_pickle", lambda p: calls.append(p) or real_read(p))
    again = tindex.load_glb(path, cache=True)
    assert calls == []                                  # served from the sidecar
    _same(ref, again)

    with open(path, "ab") as fh:                        # same index, different file: the sidecar is stale
        fh.write(b"\\n")
    third = tindex.load_glb(path, cache=True)
    assert len(calls) == 1
    _same(ref, third)
    tindex.load_glb(path, cache=True)
    assert len(calls) == 1                              # rewritten for the new file

    with open(side, "r+b") as fh:                       # damaged sidecar: ignored, rebuilt
        fh.seek(30)
        fh.write(b"\\0" * 64)
    fourth = tindex.load_glb(path, cache=True)
    assert len(calls) == 2
    _same(ref, fourth)


def test_sidecar_of_another_index_is_not_used(tmp_path):
    a, b = str(tmp_path / "a.glb"), str(tmp_path / "b.glb")
    shutil.copy(os.path.join(H.GOLD, "idx_rand_a.glb"), a)
    shutil.copy(os.path.join(H.GOLD, "idx_rand_b.glb"), b)
    tindex.load_glb(a, cache=True)
    shutil.copy(a + tindex.CACHE_SUFFIX, b + tindex.CACHE_SUFFIX)
    _same(tindex.load_glb(b, cache=True), tindex.load_glb(b, cache=False))


def test_environment_switch_and_read_only_directory(monkeypatch, tmp_path):
    path = str(tmp_path / "i.glb")
    shutil.copy(os.path.join(H.GOLD, "idx_toy.glb"), path)
    monkeypatch.setenv("TEC_INDEX_CACHE", "0")
    tindex.load_glb(path)
    assert not os.path.exists(path + tindex.CACHE_SUFFIX)
    monkeypatch.setenv("TEC_INDEX_CACHE", "1")
    monkeypatch.setattr(tindex.np, "savez", lambda *a, **k: (_ for _ in ()).throw(OSError("read-only file system")))
    tindex.load_glb(path)                               # cannot write: still loads
    assert not os.path.exists(path + tindex.CACHE_SUFFIX)
    monkeypatch.undo()
    monkeypatch.setenv("TEC_INDEX_CACHE", "1")
    tindex.load_glb(path)
    assert os.path.exists(path + tindex.CACHE_SUFFIX)
