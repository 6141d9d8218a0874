#!/usr/bin/env python
"""Selected raw metrics of an ncu report → CSV (one column per captured launch).

    ncu -i gpurun_out/X.ncu-rep --page raw --csv > /tmp/raw.csv && python profiles/summarize_ncu.py /tmp/raw.csv > profiles/X_summary.csv
"""
import csv
import sys

KEEP = ("Kernel Name", "dram__bytes", "dram__throughput", "gpu__dram_throughput", "gpu__time_duration", "launch__block_size",
        "launch__grid_size", "launch__occupancy_limit", "launch__registers_per_thread", "launch__shared_mem", "lts__t_sector_hit_rate",
        "sm__cycles_elapsed.avg.per_second", "sm__throughput", "sm__warps_active", "smsp__inst_executed.sum", "smsp__issue_active",
        "sm__inst_executed_pipe", "sm__pipe_tensor", "smsp__average_warps_issue_stalled", "syslts__t_sector_hit_rate")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    w = csv.writer(sys.stdout)
    w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(launches))])
    for c, name in enumerate(hdr):
        if any(k in name for k in KEEP) and "Not Issued" not in name:
            w.writerow([name, units[c]] + [r[c] for r in launches])


if __name__ == "__main__":
    main()
