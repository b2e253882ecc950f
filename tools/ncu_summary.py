#!/usr/bin/env python3
"""Print the metrics that matter from `ncu -i X.ncu-rep --page raw --csv` (one row per launch)."""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("smsp__inst_executed.sum", "inst"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__occupancy_limit_registers", "lim_reg"),
    ("launch__occupancy_limit_shared_mem", "lim_smem"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wf"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ld_sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "ld_req"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "st_sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "st_req"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local_ld"),
    ("smsp__sass_inst_executed_op_local_st.sum", "local_st"),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("ofdm::", "")
        parts = [name]
        for k, short in KEYS:
            if k in hdr:
                i = hdr.index(k)
                v = r[i]
                try:
                    fv = float(v.replace(",", ""))
                    v = "%.4g" % fv
                except ValueError:
                    pass
                parts.append("%s=%s%s" % (short, v, units[i] if units[i] not in ("", "%") else ""))
        print("  ".join(parts))


if __name__ == "__main__":
    main()
