#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small text file for profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_xxx.txt [note...]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__inst_executed.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio", "sm__cycles_elapsed.avg.per_second",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = " ".join(sys.argv[3:])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = [f"# ncu summary of {rep}", f"# {note}" if note else "#"]
    for d in data:
        lines.append(f"kernel: {d[hdr.index('Kernel Name')]}  grid {d[hdr.index('Grid Size')]} block {d[hdr.index('Block Size')]}")
        for k in KEYS:
            if k in hdr:
                lines.append(f"  {k:70s} {d[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
        stalls = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), d[i])
                  for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_")]
        stalls = sorted(stalls, key=lambda x: -float(x[1].replace(",", "") or 0))[:8]
        lines.append("  stalled warps per issue-active cycle, top reasons: " + ", ".join(f"{n} {float(v):.2f}" for n, v in stalls))
    open(out, "w").write("\n".join(lines) + "\n")
    try:
        print("\n".join(lines))
    except BrokenPipeError:
        pass


if __name__ == "__main__":
    main()
