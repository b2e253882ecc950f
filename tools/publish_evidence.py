#!/usr/bin/env python3
"""Copy the outputs of tools/collect_evidence.sh (gpurun_out/ev/) into profiles/ in committed form:
bench lines as they were printed, the launch list trimmed to the useful columns, the ncu --set full raw page
reduced by tools/ncu_export.py, and profiles/traffic.json (dram bytes per launch of the fused receiver)."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EV = os.path.join(ROOT, "gpurun_out", "ev")
PR = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"

for src, dst in (("bench_n64.json", "bench_n64"), ("bench_reference.json", "bench_reference"), ("bench_lte.json", "bench_lte"),
                 ("sweep_nfft.jsonl", "sweep_nfft"), ("bench_lte_streams.json", "bench_lte_streams"),
                 ("bench_sync_search.jsonl", "bench_sync_search"), ("bench_blocks.jsonl", "bench_blocks")):
    ext = os.path.splitext(src)[1]
    shutil.copy(os.path.join(EV, src), os.path.join(PR, "%s_%s%s" % (TAG, dst, ext)))
for rep in ("prof_final", "prof_sync"):                                   # raw pages of the .ncu-rep files brought back
    if os.path.exists(os.path.join(EV, rep + ".ncu-rep")) and not os.path.exists(os.path.join(EV, rep + "_raw.csv")):
        with open(os.path.join(EV, rep + "_raw.csv"), "w") as f:
            subprocess.check_call(["ncu", "-i", os.path.join(EV, rep + ".ncu-rep"), "--page", "raw", "--csv"], stdout=f)
subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "ncu_export.py"), os.path.join(EV, "prof_final_raw.csv"),
                       os.path.join(PR, "%s_ncu_full_rx_tx_n64_lte.csv" % TAG)])
if os.path.exists(os.path.join(EV, "prof_sync_raw.csv")):
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "ncu_export.py"), os.path.join(EV, "prof_sync_raw.csv"),
                           os.path.join(PR, "%s_ncu_full_sync_search.csv" % TAG)])

rows = list(csv.reader(l for l in open(os.path.join(EV, "launches_bench_n64.csv")) if l.startswith('"')))
hdr = rows[0]
idx = [hdr.index(k) for k in ("ID", "Kernel Name", "Block Size", "Grid Size", "Metric Name", "Metric Unit", "Metric Value")]
with open(os.path.join(PR, "%s_launches_bench_n64.csv" % TAG), "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        r2 = [r[i] for i in idx]
        r2[1] = r2[1][:110]
        w.writerow(r2)

# one row per kernel of the acquisition capture is enough (the launches of a case are identical)
sp = os.path.join(PR, "%s_ncu_full_sync_search.csv" % TAG)
if os.path.exists(sp):
    rows_s = list(csv.reader(open(sp)))
    seen, keep = set(), rows_s[:2]
    ki_s = rows_s[0].index("Kernel Name")
    gi_s = rows_s[0].index("Grid Size") if "Grid Size" in rows_s[0] else ki_s
    for r in rows_s[2:]:
        key = (r[ki_s], r[gi_s])
        if key not in seen:
            seen.add(key)
            keep.append(r)
    with open(sp, "w", newline="") as f:
        csv.writer(f).writerows(keep)

raw = list(csv.reader(open(os.path.join(EV, "prof_final_raw.csv"))))
h, units = raw[0], raw[1]
ki, ri, wi = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for r in raw[2:]:
    if "rx_frames_kernel" not in r[ki]:
        continue
    key = "rx_n64_16qam_1m" if "FftCfg<6" in r[ki] else ("rx_lte20_64qam" if "FftCfg<11" in r[ki] else None)
    if key is None:
        continue
    b = float(r[ri].replace(",", "")) * scale[units[ri]] + float(r[wi].replace(",", "")) * scale[units[wi]]
    acc.setdefault(key, []).append(b)
traffic = {k: int(round(sum(v) / len(v))) for k, v in acc.items()}
traffic["_source"] = ("profiles/%s_ncu_full_rx_tx_n64_lte.csv: dram__bytes_read.sum + dram__bytes_write.sum per launch of "
                      "rx_frames_kernel (mean of the captured launches; ncu --set full --clock-control none)" % TAG)
json.dump(traffic, open(os.path.join(PR, "traffic.json"), "w"), indent=1)
print(traffic)
