#!/usr/bin/env python
"""Turn `ncu --set full` captures into the committed evidence under profiles/: key metrics + top stalls
(<round>_ncu_summaries.json), DRAM traffic per launch (<round>_traffic.json) and, per report that travelled back, the
executed-instruction mix by opcode and by source line (text).
Usage: tools/summarize_profiles.py DIR [ROUND=r02]   (DIR holds raw_<workload>.csv raw pages exported on the GPU box
by tools/ncu_capture.sh and, for the workloads whose report travelled back, prof_<workload>.ncu-rep; the library in
the tree must be the build the captures were taken from)"""
import csv
import glob
import io
import json
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else 'r02'
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__shared_mem_per_block_dynamic']
mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
spath = os.path.join(root, 'profiles', f'{rnd}_ncu_summaries.json')
summ = json.load(open(spath)) if os.path.exists(spath) else {}
summ = {k: v for k, v in summ.items() if not k.startswith('final')}
traffic = {}
for rawf in sorted(glob.glob(os.path.join(src, 'raw_*.csv'))):
    w = os.path.basename(rawf)[len('raw_'):-len('.csv')]
    rep = os.path.join(src, f'prof_{w}.ncu-rep')
    raw = open(rawf).read()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    kname = vals[hdr.index('Kernel Name')]
    d = {'what': f'final {rnd} build, bench.py --workload {w}', 'kernel': kname}
    for h, u, v in zip(hdr, units, vals):
        if h in keys:
            d[h] = f'{v} {u}'.strip()
    st = [(h, float(v)) for h, v in zip(hdr, vals)
          if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
    d['top_stalls_per_issue'] = {h.replace('smsp__average_warps_issue_stalled_', '').replace(
        '_per_issue_active.ratio', ''): round(v, 3) for h, v in sorted(st, key=lambda t: -t[1])[:6]}
    # executed FP64 work per solve (thread-level SASS counters: DFMA = 2 flops) next to the algorithmic model
    bj = os.path.join(src, f'bench_{w}.json')
    if os.path.exists(bj):
        bl = json.loads(open(bj).read().strip().splitlines()[-1])
        nprob = bl['config']['problems_per_gpu']

        def per_cycle(op):  # thread-level instructions per elapsed SM cycle, summed over all sub-partitions
            k = f'smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed'
            return float(vals[hdr.index(k)]) if k in hdr else 0.0
        cycles = float(vals[hdr.index('sm__cycles_elapsed.avg')])
        ex = (2 * per_cycle('dfma') + per_cycle('dmul') + per_cycle('dadd')) * cycles  # DFMA = 2 flops
        dur = float(vals[hdr.index('gpu__time_duration.sum')])
        dur_s = dur * {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0}.get(units[hdr.index('gpu__time_duration.sum')], 1e-3)
        d['executed_fp64_flop_per_solve'] = round(ex / nprob)
        d['algorithmic_flop_per_solve'] = round(bl['roofline']['algorithmic_flops_per_launch'] / nprob)
        d['executed_fp64_tflops_under_ncu'] = round(ex / dur_s / 1e12, 2)
    summ[f'final_{w}'] = d

    def g(k):
        return float(vals[hdr.index(k)]) * mult[units[hdr.index(k)]]
    traffic[w] = {'bytes': g('dram__bytes_read.sum') + g('dram__bytes_write.sum'),
                  'read_bytes': g('dram__bytes_read.sum'), 'write_bytes': g('dram__bytes_write.sum'),
                  'unit': 'bytes per launch',
                  'source': f'ncu --set full capture of `bench.py --workload {w}` (final_{w} in {rnd}_ncu_summaries.json)'}
    import re
    m = re.search(r'smpc_solve_kernel<(\d+), (\d+), (\d+), (\d+), (\d+), (\d+)>', kname)
    a = m.groups()
    mangled = f'smpc_solve_kernelILi{a[0]}ELi{a[1]}ELi{a[2]}ELb{a[3]}ELi{a[4]}ELi{a[5]}E'  # EXACT: no other variant matches
    if os.path.exists(rep):
        lib = os.path.join(root, 'nav2_social_mpc_controller_b200', 'libsmpc.so')
        mix = subprocess.run(['python', os.path.join(root, 'tools', 'ncu_opmix.py'), rep], capture_output=True, text=True).stdout
        lines = subprocess.run(['python', os.path.join(root, 'tools', 'ncu_opmix_lines.py'), rep, mangled, lib, '45'],
                               capture_output=True, text=True).stdout
        head = (f'# {kname}\n# bench.py --workload {w}, ncu --set full --clock-control none (one launch)\n'
                f'# executed warp instructions by opcode, then by source line of csrc/smpc_device.cuh (share of all, opcode mix)\n')
        open(os.path.join(root, 'profiles', f'{rnd}_final_{w}_hotspots.txt'), 'w').write(
            head + '\n'.join(mix.splitlines()[:26]) + '\n\n' + lines)
    print(w, d['gpu__time_duration.sum'], 'issue', d['smsp__issue_active.avg.pct_of_peak_sustained_active'], 'fp64',
          d['sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'], d['top_stalls_per_issue'],
          'dram MB r/w', traffic[w]['read_bytes'] / 1e6, traffic[w]['write_bytes'] / 1e6)
json.dump(summ, open(spath, 'w'), indent=1)
json.dump(traffic, open(os.path.join(root, 'profiles', f'{rnd}_traffic.json'), 'w'), indent=1)
