"""stdin: `ncu --page raw --csv`; prints the memory-side metrics that the roofline
claims rest on (DRAM bytes, L2 hit rate, LSU / shared-memory wavefront pipe, tensor pipe)."""
import csv
import sys
KEYS = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__shared_mem_per_block_static', 'launch__occupancy_limit_warps',
        'sm__maximum_warps_per_active_cycle_pct']
rows = list(csv.reader(sys.stdin))
d = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
for k in KEYS:
  if k in d:
    print('%-86s %-12s %s' % (k, d[k][0], d[k][1]))
