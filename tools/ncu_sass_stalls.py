#!/usr/bin/env python3
"""Per-SASS-instruction view of `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda`: which instructions the
warps wait on, by stall reason, with the CUDA line each belongs to.
usage: ncu_sass_stalls.py file.csv [stall_column=stall_long_sb] [top_n=25]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
col = sys.argv[2] if len(sys.argv) > 2 else "stall_long_sb"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
rows = list(csv.reader(open(path)))
hdr, fpath, cur = None, None, None
seen = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    if r[0] != "":
        cur = (fpath, r[0], r[1].strip())
        continue
    if not r[2].startswith("0x"):
        continue
    d = dict(zip(hdr[4:], r[4:]))

    def num(k):
        try:
            return float(d.get(k, "0") or 0)
        except ValueError:
            return 0.0
    if r[2] not in seen:
        seen[r[2]] = dict(sass=r[3].strip(), line=cur, samples=num("# Samples"), inst=num("Instructions Executed"),
                          stalls={k: num(k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k})
tot = sum(v["samples"] for v in seen.values())
tin = sum(v["inst"] for v in seen.values())
mix = defaultdict(float)
for v in seen.values():
    for k, x in v["stalls"].items():
        mix[k] += x
print("instructions %.4g  samples %d" % (tin, tot))
print("stall mix:", ", ".join("%s %.1f%%" % (k[6:], 100 * x / tot) for k, x in sorted(mix.items(), key=lambda kv: -kv[1]) if x))
print("top instructions by", col)
for a, v in sorted(seen.items(), key=lambda kv: -kv[1]["stalls"].get(col, 0))[:top]:
    print("%6.2f%%  %-46s %s:%s %s" % (100 * v["stalls"].get(col, 0) / tot, v["sass"][:46], v["line"][0], v["line"][1],
                                      v["line"][2][:70]))
