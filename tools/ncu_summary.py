#!/usr/bin/env python
"""Text summary of an ncu --set full report (one block per profiled launch) for profiles/.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/rN_<name>_summary.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = [
    ('gpu__time_duration.sum', 'duration'),
    ('sm__cycles_elapsed.avg', 'SM cycles elapsed'),
    ('sm__cycles_elapsed.avg.per_second', 'SM clock'),
    ('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe (tcgen05 UTCHMMA) active % of peak'),
    ('sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active', 'legacy HMMA (mma.sync) inst % of peak'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('smsp__inst_executed.sum', 'warp instructions executed'),
    ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'FMA pipe %'),
    ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'ALU pipe %'),
    ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'XU (MUFU) pipe %'),
    ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'LSU pipe %'),
    ('dram__bytes_read.sum', 'DRAM bytes read'),
    ('dram__bytes_write.sum', 'DRAM bytes written'),
    ('dram__bytes_read.sum.per_second', 'DRAM read rate'),
    ('dram__bytes_write.sum.per_second', 'DRAM write rate'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
    ('lts__t_sectors_srcunit_tex_op_read.sum', 'L2 sectors read by SMs (32 B)'),
    ('lts__t_sectors_srcunit_tex_op_write.sum', 'L2 sectors written by SMs (32 B)'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'shared-memory bank conflicts'),
    ('launch__registers_per_thread', 'registers / thread'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic smem / block'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__cluster_size', 'cluster size'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
]
col = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    print('=' * 110)
    print(r[col['Kernel Name']][:108])
    for key, label in WANT:
        if key in col:
            print(f'  {label:52s} {r[col[key]]:>18s} {units[col[key]]}')
    stalls = []
    for h, i in col.items():
        if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued'):
            try:
                stalls.append((int(float(r[i])), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError:
                pass
    tot = sum(s for s, _ in stalls) or 1
    top = ', '.join(f'{n} {100 * s / tot:.0f}%' for s, n in sorted(stalls, reverse=True)[:7])
    print(f'  {"warp-state samples (all warps incl. idle roles)":52s} {top}')
