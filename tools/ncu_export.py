#!/usr/bin/env python3
"""Keep the columns that matter of `ncu -i X.ncu-rep --page raw --csv` (one row per captured launch) so the
summary is small enough to commit under profiles/.   usage: ncu_export.py raw.csv out.csv"""
import csv
import re
import sys

KEEP = re.compile(r"^(Kernel Name|Block Size|Grid Size|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum$|"
                  r"dram__bytes_(read|write)\.sum\.per_second|gpu__dram_throughput\.avg\.pct|lts__t_sectors_srcunit_tex_op_(read|write)\.sum$|"
                  r"lts__throughput\.avg\.pct|l1tex__throughput\.avg\.pct|l1tex__data_pipe_lsu_wavefronts\.avg\.pct|"
                  r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|"
                  r"l1tex__t_(sectors|requests)_pipe_lsu_mem_global_op_(ld|st)\.sum$|sm__throughput\.avg\.pct|"
                  r"smsp__issue_active\.avg\.pct|smsp__inst_executed\.sum$|sm__inst_executed_pipe_(alu|fma|lsu)\.avg\.pct|"
                  r"launch__registers_per_thread|launch__occupancy_limit_(registers|shared_mem)|launch__shared_mem_per_block_dynamic|"
                  r"sm__warps_active\.avg\.pct|smsp__warps_eligible\.avg\.per_cycle_active|"
                  r"smsp__average_warps_issue_stalled_(long_scoreboard|short_scoreboard|wait|barrier|mio_throttle|not_selected|math_pipe_throttle|lg_throttle)_per_issue_active|"
                  r"smsp__sass_inst_executed_op_local_(ld|st)\.sum)")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = rows[0]
    cols = [i for i, h in enumerate(hdr) if KEEP.match(h)]
    with open(sys.argv[2], "w", newline="") as f:
        w = csv.writer(f)
        for r in rows:
            w.writerow([r[i] for i in cols])


if __name__ == "__main__":
    main()
