#!/usr/bin/env python3
"""Aggregate `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda` by CUDA source line.
usage: ncu_lines.py file.csv [top_n] [launch_index]"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path)))
    agg = defaultdict(lambda: defaultdict(float))
    src = {}
    fpath = None
    hdr = None
    section = -1
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            section += 1
            continue
        if hdr is None or len(r) < len(hdr) // 2:
            continue
        d = dict(zip(hdr, r))
        # duplicate "Source" header: first is CUDA, second SASS; zip keeps the last -> recover by index
        line = r[0]
        cuda_src = r[1]
        key = (fpath, line)
        src[key] = cuda_src
        a = agg[key]

        def num(k):
            try:
                return float(d.get(k, "0") or 0)
            except ValueError:
                return 0.0
        a["inst"] += num("Instructions Executed")
        a["samples"] += num("# Samples")
        a["wf"] += num("L1 Wavefronts Shared")
        a["wf_ideal"] += num("L1 Wavefronts Shared Ideal")
        for k in ("stall_long_sb", "stall_barrier", "stall_short_sb", "stall_mio", "stall_lg", "stall_math",
                  "stall_wait", "stall_not_selected", "stall_no_inst", "stall_dispatch", "stall_selected"):
            a[k] += num(k)
        a["n_sass"] += 1
    tot_inst = sum(a["inst"] for a in agg.values())
    tot_s = sum(a["samples"] for a in agg.values())
    print("total inst %.3g  samples %d  (sections=%d; divide by sections for per-launch)" % (tot_inst, tot_s, section + 1))
    st = defaultdict(float)
    for a in agg.values():
        for k, v in a.items():
            if k.startswith("stall_"):
                st[k] += v
    print("stall mix:", {k: "%.1f%%" % (100 * v / max(tot_s, 1)) for k, v in sorted(st.items(), key=lambda kv: -kv[1])})
    print("%-16s %5s %7s %7s %6s %8s %8s  %s" % ("file", "line", "inst%", "smpl%", "sass", "wf", "wf_ideal", "source"))
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["inst"])[:top]:
        print("%-16s %5s %6.2f%% %6.2f%% %6d %8.3g %8.3g  %s" % (key[0], key[1], 100 * a["inst"] / tot_inst,
                                                               100 * a["samples"] / max(tot_s, 1), a["n_sass"], a["wf"],
                                                               a["wf_ideal"], src[key].strip()[:90]))


if __name__ == "__main__":
    main()
