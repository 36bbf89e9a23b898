#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): one block per profiled launch with the metrics the roofline
discussion uses.  Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("launch__occupancy_limit_shared_mem", "occ limit smem"),
    ("launch__occupancy_limit_registers", "occ limit regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % (alt)"),
    ("lts__t_bytes.sum", "L2 bytes"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"), ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed.sum", "warp insts"), ("smsp__inst_executed.sum", "warp insts (smsp)"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor hmma pipe %"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "tensor hmma inst %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu pipe %"),
    ("smsp__cycles_active.avg", "smsp active cycles"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct", "stall long_scoreboard %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard (ratio)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_throttle (ratio)"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle (ratio)"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier (ratio)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait (ratio)"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected (ratio)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard (ratio)"),
]


def main(path):
    # accepts an .ncu-rep, or the CSV that `ncu -i x.ncu-rep --page raw --csv` wrote on the GPU box (reports are too big to bring back)
    out = open(path).read() if path.endswith(".csv") else subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"== {r[col['Kernel Name']][:110]}  grid {r[col.get('Grid Size', 0)]} block {r[col.get('Block Size', 0)]}")
        for k, label in KEYS:
            if k in col and r[col[k]] != "":
                print(f"   {label:32s} {r[col[k]]:>16s} {units[col[k]]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
