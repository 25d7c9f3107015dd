#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) per kernel: python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__maximum_warps_per_active_cycle_pct']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('=' * 100)
    for k in want:
        if k in d:
            print(f'{k:75s} {units[hdr.index(k)]:12s} {d[k]}')
    st = sorted(((float(v), k) for k, v in d.items() if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and v), reverse=True)
    print('stalls per issue:', ', '.join(f"{k.split('issue_stalled_')[1].split('_per_')[0]}={v:.2f}" for v, k in st[:8]))
