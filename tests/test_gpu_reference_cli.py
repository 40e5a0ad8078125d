"""The reference's own command line (bin/te_count of the unmodified tree in baseline/_ref) with ONE change: the class it
instantiates, te_count.measureTE, is replaced by te_counter_b200.measureTE.  Argument parsing, genome binding, the choice
of parse_* method and the save_* call are the reference's; the output files must be the reference's golden bytes.
This is the drop-in claim of SURVEY.md 8(b) exercised from the outermost caller (te_count/bin/te_count:60-125)."""
import os
import subprocess
import sys

import pytest

import helpers as H
from bam_writer import write_bam

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "baseline", "_ref", "bin", "te_count")

DRIVER = r"""
import runpy, sys, types
root, cli = sys.argv[1], sys.argv[2]
sys.path.insert(0, root)
sys.path.insert(0, root + "/baseline/_ref")
sys.modules["pysam"] = types.ModuleType("pysam")          # imported by the reference, never called by the replacement
import te_count                                             # the reference package
import te_counter_b200
te_count.measureTE = te_counter_b200.measureTE              # the one-line swap a maintainer would make
sys.argv = [cli] + sys.argv[3:]
runpy.run_path(cli, run_name="__main__")
"""


def _run_cli(args):
    if not os.path.isfile(CLI):
        pytest.skip("baseline/_ref is not installed (run __graft_entry__.build() where /root/reference exists)")
    env = dict(os.environ, TEC_INDEX_CACHE="0")
    p = subprocess.run([sys.executable, "-c", DRIVER, ROOT, CLI] + args, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    return p.stderr


@pytest.mark.parametrize("name", ["bulk_pe_rand_a", "bulk_se_rand_b", "bulk_se_appendixA"])
def test_reference_cli_bulk(tmp_path, name):
    case = H.load_case(name)
    recs = [dict(r, name=r.get("name", "r%d" % i)) for i, r in enumerate(case["records"])]
    if any(r["end"] <= r["start"] and not r.get("flag", 0) & 4 for r in recs):
        pytest.skip("case has zero-length alignments a file cannot carry")
    bam = str(tmp_path / "in.bam")
    write_bam(bam, recs)
    out = tmp_path / "out.tsv"
    args = ["-i", bam, "-o", str(out), "-g", os.path.join(H.GOLD, case["glb"]), "-m", "custom"]
    if not case["paired"]:
        args.append("--se")
    log = _run_cli(args)
    assert out.read_text() == case["expected"]["tsv"]
    assert "Arguments:" in log                                   # the reference's main() did the talking


@pytest.mark.parametrize("name", ["sc_appendixA", "sc_appendixA_strand"])
def test_reference_cli_single_cell(tmp_path, name):
    case = H.load_case(name)
    assert case["bundle_keys"] == 10_000_000 and case["pad"] == 1000      # the stock constants: nothing to pass
    bam = str(tmp_path / "in.bam")
    write_bam(bam, [dict(r, name=r.get("name", "r%d" % i)) for i, r in enumerate(case["records"])])
    wl = tmp_path / "wl.txt"
    wl.write_text("".join(w + "\n" for w in case["whitelist"]))
    out = tmp_path / "out.tsv"
    args = ["-i", bam, "-o", str(out), "-g", os.path.join(H.GOLD, case["glb"]), "-m", "custom", "--sc", "--se",
            "-w", str(wl), "--maxcells", str(case["maxcells"])]
    if case["strand"]:
        args.append("--strand")
    _run_cli(args)
    assert out.read_text() == case["expected"]["tsv"]
