#!/usr/bin/env python3
"""Per-kernel table (launches, ms, share, DRAM bytes, GB/s) of an ncu --csv launch list taken with
--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum.
usage: tools/launch_table.py LIST.csv [FIRST_ID [LAST_ID]]"""
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, mi, vi, ii, ui = (h.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = {}
for r in rows[1:]:
    i = int(r[ii])
    if not lo <= i <= hi:
        continue
    name = r[ki].split("(")[0].replace("void ", "")[:64]
    d = per.setdefault(name, {"n": set(), "ms": 0.0, "rd": 0.0, "wr": 0.0})
    d["n"].add(i)
    v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    if r[mi].startswith("gpu__time"):
        d["ms"] += v
    elif r[mi].startswith("dram__bytes_read"):
        d["rd"] += v
    elif r[mi].startswith("dram__bytes_write"):
        d["wr"] += v
tot = sum(d["ms"] for d in per.values())
print("%d launches, %.1f ms under ncu\n" % (sum(len(d["n"]) for d in per.values()), tot))
print("| kernel | launches | ms | share | DRAM read GB | DRAM write GB | GB/s |\n|---|---|---|---|---|---|---|")
for name, d in sorted(per.items(), key=lambda kv: -kv[1]["ms"]):
    print("| `%s` | %d | %.3f | %.1f %% | %.2f | %.2f | %.0f |" % (name, len(d["n"]), d["ms"], 100 * d["ms"] / tot, d["rd"] / 1e9, d["wr"] / 1e9,
                                                              (d["rd"] + d["wr"]) / 1e6 / max(d["ms"], 1e-9)))
