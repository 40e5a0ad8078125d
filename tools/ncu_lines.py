#!/usr/bin/env python3
"""Per-source-line executed-instruction / stall-sample summary of an .ncu-rep (needs -lineinfo).
usage: tools/ncu_lines.py REPORT [TOP]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
hdr = None; rows = []
for row in csv.reader(out.splitlines()):
    if row and row[0] == 'Line No': hdr = row; continue
    if hdr and row and row[0].isdigit() and row[hdr.index('Instructions Executed')].isdigit(): rows.append(row)
ie = hdr.index('Instructions Executed'); ns = hdr.index('# Samples'); te = hdr.index('Thread Instructions Executed')
I = lambda v: int(v) if v.isdigit() else 0
tot = sum(I(x[ie]) for x in rows); tots = sum(I(x[ns]) for x in rows)
print('total warp instructions', tot, 'samples', tots)
rows.sort(key=lambda x: -I(x[ie]))
for x in rows[:top]:
    print(x[0].rjust(4), ('%.1f%%' % (100.0 * I(x[ie]) / tot)).rjust(6), ('%.1f' % (I(x[te]) / max(1, I(x[ie])))).rjust(5),
          ('%.1f%%' % (100.0 * I(x[ns]) / max(1, tots))).rjust(6), x[1].strip()[:120])
