"""Prints the metrics the roofline claims rest on from an .ncu-rep file.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-index]
"""
import csv
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'launch__registers_per_thread',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
    'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum',
    'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
    'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
    'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
]


def main():
  rep = sys.argv[1]
  idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
  out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'],
                       capture_output=True, text=True).stdout
  rows = list(csv.reader(out.splitlines()))
  hdr, units, vals = rows[0], rows[1], rows[2 + idx]
  d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
  print('kernel:', d.get('Kernel Name', ('', ''))[1][:120])
  for k in KEYS:
    if k in d:
      print('%-86s %-12s %s' % (k, d[k][0], d[k][1]))


if __name__ == '__main__':
  main()
