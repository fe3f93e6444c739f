#!/usr/bin/env python3
"""Summarise an ncu --set full report into profiles/: one CSV row per kernel and
the DRAM traffic of the IJ apply GEMM (bench.py's roofline.traffic).

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/ncu_full_rNN_kernels.csv "<how it was captured>"
"""
import csv
import json
import os
import subprocess
import sys

rep, out_csv, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else '')
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
COLS = [('gpu__time_duration.sum', 'duration'), ('dram__bytes_read.sum', 'dram_read'),
        ('dram__bytes_write.sum', 'dram_write'),
        ('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'tensor_pipe_pct'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
        ('lts__t_sector_hit_rate.pct', 'l2_hit_pct'), ('launch__registers_per_thread', 'regs'),
        ('launch__occupancy_limit_shared_mem', 'occ_limit_smem'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_pct'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_active_pct'),
        ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall_math_throttle'),
        ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall_wait'),
        ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall_short_sb'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall_long_sb'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall_barrier'),
        ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'stall_lg_throttle'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem_bank_conflicts')]
COLS += [('sm__inst_executed_pipe_tc.sum', 'inst_pipe_tc'), ('sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tc_pipe_pct'),
         ('lts__t_bytes.sum', 'l2_bytes'), ('l1tex__m_xbar2l1tex_read_bytes.sum', 'l2_to_sm_bytes')]
COLS = [(c, n) for c, n in COLS if c in idx]
SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
best = {}
for r in data:
    if 'nan' in r[idx['gpu__time_duration.sum']].lower() or 'nan' in r[idx['dram__bytes_read.sum']].lower():
        continue                      # ncu sometimes returns NaN for a replayed launch: keep a complete one
    best[r[idx['Kernel Name']]] = r   # last complete launch of each kernel (warm)
lines = ['# ' + note, 'kernel,' + ','.join('{} [{}]'.format(n, units[idx[c]]) for c, n in COLS)]
traffic = {}
for name, r in best.items():
    vals = []
    for c, n in COLS:
        try:
            vals.append('%.5g' % float(r[idx[c]]))
        except ValueError:
            vals.append(r[idx[c]])
    lines.append('"' + name.replace('"', "'")[:110] + '",' + ','.join(vals))
    if 'dgemm_kernel<0, 0, 0' in name:
        rd = float(r[idx['dram__bytes_read.sum']]) * SCALE[units[idx['dram__bytes_read.sum']]]
        wr = float(r[idx['dram__bytes_write.sum']]) * SCALE[units[idx['dram__bytes_write.sum']]]
        traffic = {'ij_apply_dram_bytes_per_launch_capture': rd + wr, 'capture': note, 'report_csv': out_csv}
open(out_csv, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines))
if traffic:
    tpath = os.path.join(os.path.dirname(out_csv), 'ncu_traffic.json')
    n_capture = float(os.environ.get('VT_CAPTURE_N', '1e6'))
    traffic['ij_apply_dram_bytes_per_obs'] = traffic['ij_apply_dram_bytes_per_launch_capture'] / n_capture
    traffic['algorithmic_bytes_per_obs'] = 16 * 1024
    json.dump(traffic, open(tpath, 'w'), indent=1)
    print(traffic)
