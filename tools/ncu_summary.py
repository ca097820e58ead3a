#!/usr/bin/env python
"""Print the headline metrics of an .ncu-rep (read here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [launch-id]"""
import csv
import subprocess
import sys

WANT = ['Duration', 'DRAM Throughput', 'Memory Throughput', 'L2 Hit Rate',
        'Executed Ipc Active', 'Issue Slots Busy', 'Registers Per Thread',
        'Theoretical Occupancy', 'Achieved Occupancy', 'Executed Instructions',
        'Eligible Warps Per Scheduler', 'No Eligible',
        'Warp Cycles Per Issued Instruction', 'L1/TEX Hit Rate', 'Mem Busy',
        'Dynamic Shared Memory Per Block', 'Block Limit Shared Mem',
        'Block Limit Registers', 'Waves Per SM', 'Grid Size', 'Block Size',
        'Compute (SM) Throughput', 'Mem Pipes Busy', 'Local Load Instructions',
        'Local Store Instructions']
RAW = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum',
       'smsp__inst_executed.sum',
       'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
       'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']


def page(rep, name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'],
                         capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep = sys.argv[1]
    lid = sys.argv[2] if len(sys.argv) > 2 else '0'
    rows = page(rep, 'details')
    idx = {h: i for i, h in enumerate(rows[0])}
    for r in rows[1:]:
        if r[idx['ID']] == lid and r[idx['Metric Name']] in WANT:
            print('%-34s %s %s' % (r[idx['Metric Name']], r[idx['Metric Value']],
                                   r[idx['Metric Unit']]))
        if r[idx['ID']] == lid and r[idx['Metric Name']] == 'Duration':
            print('kernel:', r[idx['Kernel Name']][:110])
    rows = page(rep, 'raw')
    hdr = rows[0]
    for i, h in enumerate(hdr):
        if h in RAW or 'pcsamp_warps_issue_stalled' in h:
            vals = [r[i] for r in rows[1:]]
            print('%-64s %s' % (h, vals))


if __name__ == '__main__':
    main()
