#!/usr/bin/env python3
"""Print the headline metrics of the kernels in an .ncu-rep (run in the build container; reads `ncu -i ... --page raw --csv`).
usage: python tools/ncu_summary.py report.ncu-rep [kernel-name-substring]"""
import csv
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(gpu__time_duration.sum|sm__throughput.avg.pct_of_peak_sustained_elapsed|dram__throughput.avg.pct_of_peak_sustained_elapsed|"
    r"sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__block_size|"
    r"smsp__inst_executed.sum|smsp__issue_active.avg.pct_of_peak_sustained_active|dram__bytes_read.sum|dram__bytes_write.sum|"
    r"lts__t_sector_hit_rate.pct|l1tex__t_sector_hit_rate.pct|sm__inst_executed_pipe_(alu|fma|fmaheavy|lsu|uniform).avg.pct_of_peak_sustained_active|"
    r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio|sass__inst_executed_local_(loads|stores)|smsp__thread_inst_executed.sum)$")


def main():
    rep = sys.argv[1]
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if flt not in name:
            continue
        print("## " + name[:120])
        for i, c in enumerate(hdr):
            if KEEP.match(c):
                v = r[i]
                if "stalled" in c:
                    try:
                        if float(v.replace(",", "")) < 0.15:
                            continue
                    except ValueError:
                        pass
                print(f"{c:95s} {v} {units[i]}")


if __name__ == "__main__":
    main()
