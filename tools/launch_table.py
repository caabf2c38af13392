"""Per-launch table (time, DRAM bytes) from an `ncu --csv` launch list: python tools/launch_table.py file.csv [first|last]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui, mi, gi = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Unit", "Metric Name", "Grid Size"))
byid = collections.OrderedDict()
for r in rows[1:]:
    d = byid.setdefault(r[0], {"k": re.sub(r"\(.*", "", r[ki]), "g": r[gi]})
    d[r[mi]] = float(r[vi].replace(",", ""))
    d[r[mi] + "_u"] = r[ui]
L = [d for d in byid.values() if "at::" not in d["k"] and "elementwise" not in d["k"]]
n = int(sys.argv[2]) if len(sys.argv) > 2 else len(L)
tot = 0.0
for d in L[:n]:
    t = d["gpu__time_duration.sum"]
    t = t / 1e3 if d["gpu__time_duration.sum_u"] == "ns" else t
    rd, wr = d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0)
    tot += t
    print(f"{d['k'].replace('void hn::', ''):36s} {d['g']:14s} {t:8.1f} us  rd {rd / 1e6:8.1f} MB wr {wr / 1e6:8.1f} MB {(rd + wr) / t / 1e3:6.0f} GB/s")
print(f"total {tot:.1f} us over {min(n, len(L))} launches")
