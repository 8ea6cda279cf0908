#!/usr/bin/env python
"""Print the metrics that matter for the solve kernel from an ncu raw page (.ncu-rep or the raw_*.csv exported by
tools/ncu_capture.sh): duration, occupancy, issue / pipe utilisation, local-memory traffic, top stall reasons."""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__shared_mem_per_block_dynamic',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum']


def load(path):
    if path.endswith('.csv'):
        raw = open(path).read()
    else:
        raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    return rows[0], rows[1], rows[2]


for rep in sys.argv[1:]:
    hdr, units, vals = load(rep)
    print('==', rep, vals[hdr.index('Kernel Name')][:70])
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            print('  ', h, v, u)
    st = [(h, float(v)) for h, v in zip(hdr, vals)
          if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
    for h, v in sorted(st, key=lambda t: -t[1])[:9]:
        print('   stall', h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''),
              round(v, 3))
