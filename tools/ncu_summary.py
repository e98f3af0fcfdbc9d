#!/usr/bin/env python
"""Print the key metrics of an .ncu-rep (first kernel) — used to write profiles/*.md summaries."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active',
        'sm__pipe_tensor', 'sm__inst_executed_pipe_tensor', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit', 'sm__cycles_elapsed.avg',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_sector_hit_rate.pct', 'launch__grid_size',
        'launch__block_size', 'smsp__average_warp', 'smsp__warp_issue_stalled', 'launch__waves_per_multiprocessor',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared', 'sm__inst_executed_pipe_lsu', 'smsp__inst_executed_pipe_alu',
        'smsp__inst_executed_pipe_fma', 'sm__pipe_tensor_cycles_active', 'sm__inst_executed_pipe_uniform']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    print('==', vals[hdr.index('Kernel Name')][:90] if 'Kernel Name' in hdr else '')
    for i, h in enumerate(hdr):
        if any(h.startswith(w) for w in WANT) and 'per_second' not in h:
            v = vals[i]
            if 'stalled' in h:
                try:
                    if float(v) < 0.05:
                        continue
                except ValueError:
                    pass
            print(f'  {h} [{units[i]}] = {v}')
