#!/usr/bin/env python3
"""Print the handful of raw ncu metrics we track, for one report:  python tools/ncu_summary.py prof.ncu-rep [kernel-row]"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'gpu__dram_throughput.avg.pct',
        'sm__cycles_active.avg', 'sm__cycles_active.m', 'sm__cycles_elapsed.avg ', 'smsp__inst_executed.sum ',
        'smsp__issue_active.avg.pct', 'smsp__average_warps_issue_stalled',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ', 'smsp__warps_eligible.avg', 'sm__warps_active.avg.per_cycle_active',
        'launch__registers_per_thread', 'launch__grid_size', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum ',
        'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed_op_shared_ld.sum', 'sm__inst_executed_pipe_lsu.sum ']


def main():
    rep = sys.argv[1]
    row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2 + row]
    for h, u, v in zip(hdr, units, vals):
        if h in ("Kernel Name",) or any(h.startswith(w.strip()) for w in WANT):
            print(f"{h:90s} {u:12s} {v}")


if __name__ == "__main__":
    main()
