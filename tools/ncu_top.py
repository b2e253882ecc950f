#!/usr/bin/env python3
"""Top CUDA source lines by executed instructions (and stall samples) for one kernel launch.
usage: ncu_top.py report.ncu-rep kernel_regex [launch_skip] [top_n]"""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "-k",
                          "regex:" + rx, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fpath = hdr = None
    agg, smp, src = defaultdict(float), defaultdict(float), {}
    stall = defaultdict(lambda: defaultdict(float))
    skeys = ["stall_long_sb", "stall_wait", "stall_short_sb", "stall_barrier", "stall_not_selected", "stall_selected",
             "stall_math", "stall_mio", "stall_lg", "stall_no_inst", "stall_dispatch", "stall_branch_resolving"]
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            print(r[1][:120])
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r[0].strip():
            continue
        try:
            k = (fpath, int(r[0]))
            agg[k] += float(r[hdr.index("Instructions Executed")] or 0)
            smp[k] += float(r[hdr.index("# Samples")] or 0)
            for s in skeys:
                if s in hdr:
                    stall[s][k] += float(r[hdr.index(s)] or 0)
        except ValueError:
            continue
        src[k] = r[1]
    tot, ts = sum(agg.values()), sum(smp.values())
    print("total warp-inst %.4g   samples %d" % (tot, ts))
    print("stall mix: " + "  ".join("%s %.1f%%" % (s.replace("stall_", ""), 100 * sum(stall[s].values()) / max(ts, 1))
                                     for s in sorted(skeys, key=lambda s: -sum(stall[s].values()))[:8]))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
        worst = max(skeys, key=lambda s: stall[s][k])
        print("%-14s %4d inst %5.1f%% smp %5.1f%% (%s)  %s" % (k[0], k[1], 100 * v / tot, 100 * smp[k] / max(ts, 1),
                                                             worst.replace("stall_", ""), src[k].strip()[:95]))


if __name__ == "__main__":
    main()
