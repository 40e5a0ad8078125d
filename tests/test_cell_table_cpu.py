"""The cell table builder (te_counter_b200/csrc/stab_build.h, plain C++) against brute force on the
CPU: every position of random small indices, all supported cell sizes, overflow chains, empty
intervals, and the STAB_EXT extension that lets one sector answer both points of a pair across a
cell border.  The CUDA kernels decode exactly what this host-side reader decodes."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cell_table_against_brute_force(tmp_path):
    exe = str(tmp_path / "stab_selftest")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tools", "stab_selftest.cpp")], check=True)
    r = subprocess.run([exe, "8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "self-test ok" in r.stdout


def test_cell_table_layout2_against_brute_force(tmp_path):
    """Layout 2 (stab2_build.h) behind the default two-pass bulk kernels: thresholds, twin pairs, EDGE cells,
    overflow chains -- every unit of random indices against brute force through the scalar reader; every deferred unit
    also through the scalar statement of the two-sector kernel (bulk2_pair_kernel)."""
    exe = str(tmp_path / "stab2_selftest")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tools", "stab2_selftest.cpp")], check=True)
    r = subprocess.run([exe, "9"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "self-test ok" in r.stdout
