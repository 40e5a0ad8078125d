"""te_counter_b200/synth_bam.py (the file form of the synthetic workload, used by bench.py's from_file
leg and tools/file_e2e.py): a valid BAM that the Python reader, libtecbam and the block-parallel
decoder read identically."""
import numpy as np
import pytest

from te_counter_b200 import bam, reads, synth_bam
from test_bgzf_dev_cpu import _decode, _native, lib       # noqa: F401  (lib is a fixture)

KEYS = [str(i) for i in range(1, 23)] + ["X", "Y", "M"]


@pytest.mark.parametrize("mode", ["se", "pe", "sc"])
def test_synthetic_bam_is_read_identically(lib, tmp_path, mode):      # noqa: F811
    path, wlf = str(tmp_path / "s.bam"), str(tmp_path / "wl.txt")
    n = 6000
    size = synth_bam.write(path, n, mode, whitelist=wlf, n_barcodes=500, procs=2)
    assert size > 100000
    f = bam.AlignmentFile(path, "r")
    assert f.references[:3] == ["chr1", "chr2", "chr3"] and len(f.references) == 25
    recs = list(f)
    f.close()
    assert len(recs) == n
    assert all(r.reference_end is None or r.reference_end >= r.reference_start + 100 for r in recs)
    if mode == "sc":
        tags = dict(recs[0].get_tags())
        assert len(tags["CB"]) == 18 and tags["CB"].endswith("-1") and len(tags["UB"]) == 12
    wl = reads.Whitelist(wlf) if mode == "sc" else None
    want = _native(path, mode, reads.ChromMap(KEYS), wl, 20)
    assert len(want["start"]) == n
    assert want["start"].tolist() == [r.reference_start for r in recs] or mode == "sc"
    if mode == "sc":
        assert (want["cell"] != 0xFFFFFFFF).mean() > 0.8            # most barcodes are on the whitelist
    rc, msg, got = _decode(lib, path, mode, reads.ChromMap(KEYS), wl, 20, 7)
    assert rc == 0, msg
    for k in want:
        assert np.array_equal(want[k], got[k]), k
