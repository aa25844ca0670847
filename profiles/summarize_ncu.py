#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the few counters the roofline discussion uses.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>_full.csv

One row per profiled launch; runs here (no GPU needed: it only calls `ncu -i ... --page raw --csv`).
"""
import csv
import subprocess
import sys

KEEP = [
    ('Kernel Name', 'kernel'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__shared_mem_per_block_static', 'smem_static_B'),
    ('gpu__time_duration.sum', 'duration'),
    ('smsp__inst_executed.sum', 'warp_instructions'),
    ('smsp__thread_inst_executed_per_inst_executed.ratio', 'active_lanes_per_inst'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_active_pct'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_pct'),
    ('sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active', 'pipe_fp64_pct'),
    ('sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active', 'pipe_fma_pct'),
    ('sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active', 'pipe_alu_pct'),
    ('sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active', 'pipe_lsu_pct'),
    ('sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active', 'pipe_xu_pct'),
    ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall_barrier'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall_long_scoreboard'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall_short_scoreboard'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall_wait'),
    ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall_math_pipe'),
    ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'stall_lg_throttle'),
    ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'stall_not_selected'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem_bank_conflicts'),
    ('l1tex__t_sector_hit_rate.pct', 'l1_hit_pct'),
    ('lts__t_sector_hit_rate.pct', 'l2_hit_pct'),
    ('dram__bytes_read.sum', 'dram_read'),
    ('dram__bytes_write.sum', 'dram_write'),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    w = csv.writer(sys.stdout)
    w.writerow([name + (f' [{units[col[k]]}]' if k in col and units[col[k]] else '') for k, name in KEEP])
    for r in rows[2:]:
        w.writerow([r[col[k]] if k in col else '' for k, _ in KEEP])


if __name__ == '__main__':
    main()
