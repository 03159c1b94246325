#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here on the CPU box with `ncu -i ... --page raw --csv`) into a small JSON
under profiles/: duration, DRAM bytes, pipe utilisation, issue rate, warp-execution efficiency, registers."""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed": "pipe_fmaheavy_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "pipe_tensor_pct",
    "smsp__issue_active.avg.per_cycle_active": "issue_per_cycle_per_smsp",
    "smsp__inst_executed.sum": "warp_insts",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_warp_inst",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "lts__t_bytes.sum": "l2_bytes",
    "smsp__cycles_elapsed.avg.per_second": "sm_clock",
    "smsp__cycles_elapsed.max": "sm_cycles",
    "lts__t_sectors_srcunit_tex.sum": "l2_sectors_from_sm",
}


def main():
    rep = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for h, u, v in zip(hdr, units, r):
            if h in KEYS:
                try:
                    d[KEYS[h]] = float(v.replace(",", ""))
                except ValueError:
                    d[KEYS[h]] = v
                d[KEYS[h] + "_unit"] = u
        res.append(d)
    s = json.dumps(res, indent=1)
    if out:
        with open(out, "w") as f:
            f.write(s + "\n")
    print(s)


if __name__ == "__main__":
    main()
