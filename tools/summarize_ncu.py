#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep and the launch list into the text summaries kept under profiles/."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
STALL = "smsp__average_warps_issue_stalled_"


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    return dict(zip(r[0], zip(r[1], r[2])))


def main():
    for rep in sys.argv[1:]:
        d = raw(rep)
        print(f"## {rep}\n")
        print(f"kernel: {d.get('Kernel Name', ('', '?'))[1]}\n")
        for k in KEYS:
            if k in d:
                print(f"{k:75s} {d[k][1]:>18s} {d[k][0]}")
        st = sorted(((float(v[1]), k[len(STALL):].replace('_per_issue_active.ratio', '')) for k, v in d.items()
                     if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v[1] not in ("", "n/a")), reverse=True)
        print("\nwarp stall cycles per issued instruction: " + ", ".join(f"{n} {x:.2f}" for x, n in st[:8]) + "\n")


if __name__ == "__main__":
    main()
