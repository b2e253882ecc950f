#!/usr/bin/env python3
"""Shared-memory wavefronts per SASS instruction from `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda`:
where the L1 data pipe's work goes and which instructions exceed their conflict-free ideal.
usage: ncu_wavefronts.py file.csv [top_n=12]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
hdr, seen, cur, fp = None, {}, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fp = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    if r[0] != "":
        cur = (fp, r[0], r[1].strip())
        continue
    if r[2].startswith("0x") and r[2] not in seen:
        d = dict(zip(hdr[4:], r[4:]))

        def num(k):
            try:
                return float(d.get(k, "0") or 0)
            except ValueError:
                return 0.0
        seen[r[2]] = (r[3].strip(), cur, num("L1 Wavefronts Shared"), num("L1 Wavefronts Shared Ideal"))
tot = sum(v[2] for v in seen.values())
ideal = sum(v[3] for v in seen.values())
print("shared-memory wavefronts %.4g, conflict-free ideal %.4g (%.0f %% excess)" % (tot, ideal, 100 * (tot - ideal) / max(ideal, 1)))
print("largest excess:")
for a, v in sorted(seen.items(), key=lambda kv: -(kv[1][2] - kv[1][3]))[:top]:
    print("%9.3g over %9.3g  %-44s %s:%s %s" % (v[2] - v[3], v[3], v[0][:44], v[1][0], v[1][1], v[1][2][:60]))
