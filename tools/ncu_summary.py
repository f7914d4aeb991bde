#!/usr/bin/env python
"""Summarise an .ncu-rep: per-kernel headline metrics (raw page) and an opcode histogram with
stall samples (source page, SASS view).  Usage: tools/ncu_summary.py prof.ncu-rep [out.md]"""
import collections
import csv
import io
import re
import subprocess
import sys

RAW = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
       'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
       'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
       'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
       'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
       'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
       'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
       'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
       'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
       'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
       'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
       'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active']


def ncu(rep, page, extra=()):
    return subprocess.run(['ncu', '-i', rep, '--page', page, '--csv', *extra], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    out = open(sys.argv[2], 'w') if len(sys.argv) > 2 else sys.stdout
    rows = list(csv.reader(io.StringIO(ncu(rep, 'raw'))))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f'# ncu summary of {rep}\n', file=out)
    for r in rows[2:]:
        print(f"## {r[col['Kernel Name']][:90]}  (id {r[col['ID']]})\n", file=out)
        for m in RAW:
            if m in col:
                print(f'- `{m}` = {r[col[m]]} {units[col[m]]}', file=out)
        print(file=out)
    # SASS histogram per kernel
    text = ncu(rep, 'source')
    kernel, hist, samples, stalls = None, None, None, None
    cols = None

    seen = set()

    def flush():
        if kernel is None or not hist or kernel in seen:
            return
        seen.add(kernel)
        tot = sum(hist.values())
        print(f'## SASS mix: {kernel[:90]}\n', file=out)
        print(f'total warp instructions {tot}; stall samples {sum(samples.values())}\n', file=out)
        print('| opcode | warp instr | % | stall samples | % |', file=out)
        print('|---|---|---|---|---|', file=out)
        ts = max(sum(samples.values()), 1)
        for op, n in sorted(hist.items(), key=lambda kv: -kv[1])[:22]:
            print(f'| {op} | {n} | {100 * n / tot:.1f} | {samples[op]} | {100 * samples[op] / ts:.1f} |', file=out)
        print('\nstall reasons (all samples): ' +
              ', '.join(f'{k}={v}' for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]) + '\n', file=out)

    for row in csv.reader(io.StringIO(text)):
        if not row:
            continue
        if row[0] == 'Kernel Name':
            flush()
            kernel, hist, samples, stalls = row[1], collections.Counter(), collections.Counter(), collections.Counter()
            cols = None
            continue
        if row[0] == 'Address':
            cols = {h: i for i, h in enumerate(row)}
            continue
        if cols is None or kernel is None or not row[0].startswith('0x'):
            continue
        sass = row[cols['Source']].strip()
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', sass)
        if not m:
            continue
        op = m.group(2)
        n = int(row[cols['Instructions Executed']] or 0)
        s = int(row[cols['Warp Stall Sampling (All Samples)']] or 0)
        hist[op] += n
        samples[op] += s
        for k, i in cols.items():
            if k.startswith('stall_') and 'Not Issued' not in k:
                try:
                    stalls[k] += int(row[i] or 0)
                except ValueError:
                    pass
    flush()


if __name__ == '__main__':
    main()
