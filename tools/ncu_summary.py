#!/usr/bin/env python3
"""Summarise an .ncu-rep: per-kernel time, DRAM bytes, throughput percentages, occupancy.
usage: ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "ipc"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_branch"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
        rows = list(csv.reader(io.StringIO(out)))
        rows = [r for r in rows if len(r) > 10]
        hdr, units = rows[0], rows[1]
        kn = hdr.index("Kernel Name")
        print("== %s" % path)
        for r in rows[2:]:
            print("-- %s" % r[kn][:70])
            parts = []
            for key, short in WANT:
                if key in hdr:
                    i = hdr.index(key)
                    parts.append("%s=%s%s" % (short, r[i], (" " + units[i]) if units[i] and short in ("time", "dram_rd", "dram_wr", "l2_bytes") else ""))
            print("   " + "  ".join(parts))


if __name__ == "__main__":
    main()
