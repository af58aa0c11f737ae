#!/usr/bin/env python
"""Print the key metrics of an .ncu-rep (first kernel): `python profiles/ncu_summary.py rep.ncu-rep`."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__lsu_writeback_active.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def main(path, idx=0):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    vals = rows[2 + idx]
    print("kernel:", vals[hdr.index("Kernel Name")])
    for i, h in enumerate(hdr):
        if h in KEYS or h.startswith("smsp__average_warps_issue_stalled"):
            if h.startswith("smsp__average_warps_issue_stalled") and float(vals[i] or 0) < 0.3:
                continue
            print("  %-78s %s %s" % (h, vals[i], units[i]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
