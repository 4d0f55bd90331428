#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep (ncu --set full) into a small text file for profiles/.
usage: tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/name.txt ["title"]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__warps_eligible.avg.per_cycle_active",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["# " + title, "# source: ncu --set full --clock-control none (one launch); values per launch", ""]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        lines.append("kernel: %s" % d.get("Kernel Name", ("", "?"))[1])
        for k in KEYS:
            if k in d:
                lines.append("  %-70s %s %s" % (k, d[k][1], d[k][0]))
        st = sorted(((float(v[1] or 0), k[len(STALLS):].replace("_per_issue_active.ratio", "")) for k, v in d.items()
                     if k.startswith(STALLS) and k.endswith("_per_issue_active.ratio")), reverse=True)
        lines.append("  stall reasons (warps stalled per issue-active cycle):")
        for v, k in st[:8]:
            lines.append("    %-30s %.3f" % (k, v))
        lines.append("")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
