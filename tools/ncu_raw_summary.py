#!/usr/bin/env python
"""Key metrics + stall reasons from `ncu -i X.ncu-rep --page raw --csv` (stdin)."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__cycles_elapsed.avg', 'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print("%-70s %s %s" % (w, vals[i], units[i]))
st = [(h.replace('smsp__pcsamp_warps_issue_stalled_', ''), int(float(vals[i] or 0))) for i, h in enumerate(hdr)
      if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h]
tot = sum(v for _, v in st) or 1
print("stalls: " + ", ".join("%s %.1f%%" % (n, 100.0 * v / tot) for n, v in sorted(st, key=lambda kv: -kv[1]) if v * 100 >= tot))
