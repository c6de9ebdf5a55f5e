#!/usr/bin/env python
"""Summarise an .ncu-rep (from `ncu --set full`) into a short text file under
profiles/: duration, issue-slot / pipe utilisation, occupancy, DRAM traffic,
stall reasons.  Usage: tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_static",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# summary of {rep} (ncu --set full --clock-control none); one block per captured launch"]
    for vals in rows[2:]:
        d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
        lines.append("")
        for k in KEYS:
            if k in d:
                lines.append(f"{k:80s} {d[k][0]} {d[k][1]}")
        stalls = sorted(((float(v[0]), k[len(STALLS):].replace("_per_issue_active.ratio", "")) for k, v in d.items()
                         if k.startswith(STALLS) and k.endswith("_per_issue_active.ratio") and v[0]), reverse=True)
        lines.append("warp stall reasons (warps per issue-active cycle): " + ", ".join(f"{n} {x:.2f}" for x, n in stalls[:8]))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
