"""One BAM file counted by two ranks through measureTE (te_counter_b200/shard.py, tec_bam_count_range): the TSV files must
be byte for byte those of the one-rank run.  The two processes share GPU 0 and talk over gloo (host-side all-reduce,
callback collectives); with two GPUs on the box the same is launched over NCCL, where the library issues the collectives."""
import os
import subprocess
import sys

import pytest

import helpers as H
from bam_writer import write_bam
import numpy as np

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _records(n, seed, sc):
    """Coordinate-free random records over the chromosomes of idx_rand_a.glb (features up to 120 kbp), mate names without
    '/', CB / UB tags for single cell (60 barcodes of which 50 are whitelisted, a few hundred UMIs)."""
    rng = np.random.default_rng(seed)
    wl = ["ACGTAC%04d" % i for i in range(50)]
    umis = ["".join(rng.choice(list("ACGT"), size=10)) for _ in range(300)]
    recs = []
    for i in range(n):
        start = int(rng.integers(0, 121000))
        flag = 0
        for bit, p in ((0x4, 0.02), (0x10, 0.5), (0x200, 0.02), (0x400, 0.03)):
            if rng.random() < p:
                flag |= bit
        r = {"chrom": ["chr1", "chr2", "chrX", "chr7"][int(rng.integers(4))], "start": start, "end": start + int(rng.choice([30, 75, 100, 4000])),
             "flag": flag, "mapq": int(rng.choice([0, 19, 20, 60, 255])), "name": "q%d" % (i // 2)}
        if sc:
            r["CB"] = wl[int(rng.integers(50))] if rng.random() < 0.9 else "TTTTTT%04d" % int(rng.integers(10))
            r["UB"] = umis[int(rng.integers(300))]
        recs.append(r)
    if sc:
        recs.sort(key=lambda r: (r["chrom"], r["start"]))
    return recs, wl


def _run(world, port, args, extra_env=None):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), TEC_INDEX_CACHE="0")
    env.update(extra_env or {})
    if world == 1:
        cmd = [sys.executable, "-m", "te_counter_b200.sharded"] + args
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
               "--master-port", str(port), "-m", "te_counter_b200.sharded"] + args
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    return p.stderr


def _worlds():
    return [2]


def test_bulk_se_file_split_over_ranks(tmp_path):
    recs, _ = _records(30000, 77, False)
    bam = str(tmp_path / "se.bam")
    write_bam(bam, recs, block=20000)
    glb = os.path.join(H.GOLD, "idx_rand_a.glb")
    one = str(tmp_path / "one.tsv")
    _run(1, 0, ["--glb", glb, "--bam", bam, "--mode", "se", "-o", one])
    for world in _worlds():
        out = str(tmp_path / ("w%d.tsv" % world))
        log = _run(world, 29600 + world, ["--glb", glb, "--bam", bam, "--mode", "se", "-o", out])
        assert open(out, "rb").read() == open(one, "rb").read()
        assert "decoded by this rank" in log                                  # the ranges were used, not the fallback
    # paired end stays with one decoder; the merged result is the same
    recs, _ = _records(20001, 78, False)
    bam = str(tmp_path / "pe.bam")
    write_bam(bam, recs, block=20000)
    one = str(tmp_path / "one_pe.tsv")
    _run(1, 0, ["--glb", glb, "--bam", bam, "--mode", "pe", "-o", one])
    out = str(tmp_path / "w2_pe.tsv")
    _run(2, 29610, ["--glb", glb, "--bam", bam, "--mode", "pe", "-o", out])
    assert open(out, "rb").read() == open(one, "rb").read()


def test_single_cell_file_split_over_ranks(tmp_path):
    recs, wl = _records(30000, 79, True)
    bam = str(tmp_path / "sc.bam")
    write_bam(bam, recs, block=20000)
    wlf = tmp_path / "wl.txt"
    wlf.write_text("".join(w + "\n" for w in wl))
    glb = os.path.join(H.GOLD, "idx_rand_a.glb")
    base = ["--glb", glb, "--bam", bam, "--mode", "sc", "--whitelist", str(wlf), "--maxcells", "40", "--strand"]
    one = str(tmp_path / "one.tsv")
    _run(1, 0, base + ["-o", one])
    for world in _worlds():
        out = str(tmp_path / ("w%d.tsv" % world))
        _run(world, 29620 + world, base + ["-o", out])
        assert open(out, "rb").read() == open(one, "rb").read()
        assert open(out.replace(".tsv", ".barcode_freq.tsv"), "rb").read() == open(one.replace(".tsv", ".barcode_freq.tsv"), "rb").read()
    # a damaged range start on one rank: the ranks fall back to one decoder together and the result stands
    out = str(tmp_path / "fallback.tsv")
    _run(2, 29630, base + ["-o", out], {"TEC_TEST_SHARD_REFUSE": "1"})
    assert open(out, "rb").read() == open(one, "rb").read()
