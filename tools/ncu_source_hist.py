"""Executed warp instructions by opcode (and the hottest SASS lines) from an ncu report's source page.
usage: python tools/ncu_source_hist.py report.ncu-rep kernel-regex [top]"""
import csv
import re
import subprocess
import sys
from collections import Counter


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{rx}'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = None
    ops, stall = Counter(), Counter()
    lines = []
    for r in rows:
        if r and r[0] == 'Address':
            if h is not None:
                break                     # first profiled launch only
            h = r
            continue
        if h is None or len(r) < len(h):
            continue
        src = r[h.index('Source')].strip()
        n = int(r[h.index('Instructions Executed')] or 0)
        smp = int(r[h.index('Warp Stall Sampling (All Samples)')] or 0)
        op = re.sub(r'^@!?U?P\d+\s+', '', src).split()[0].split('.')[0]
        ops[op] += n
        stall[op] += smp
        lines.append((n, smp, src))
    tot = sum(ops.values())
    print(f'total warp instructions {tot}')
    for op, n in ops.most_common(25):
        print(f'  {op:10s} {n:12d} {100.0 * n / tot:5.1f}%   stall samples {stall[op]}')
    if top:
        print('hottest lines by stall samples:')
        for n, smp, src in sorted(lines, key=lambda t: -t[1])[:top]:
            print(f'  {smp:7d} {n:10d}  {src}')


if __name__ == '__main__':
    main()
