#!/usr/bin/env python
"""Per-kernel time share from an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
usage: python tools/launch_share.py profiles/r01_fused_launches.csv"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = [r for r in open(path) if r.startswith('"')]
    rd = csv.DictReader(rows)
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rd:
        if r['Metric Name'] != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(unit, 1)
        name = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '')
        name = re.sub(r'<.*', '', name)
        tot[name] += ns
        cnt[name] += 1
    all_ns = sum(tot.values())
    print(f'| kernel | launches | total ms | mean us | share |\n|---|---|---|---|---|')
    for k in sorted(tot, key=tot.get, reverse=True):
        print(f'| `{k}` | {cnt[k]} | {tot[k] / 1e6:.3f} | {tot[k] / cnt[k] / 1e3:.1f} | {100 * tot[k] / all_ns:.1f} % |')


if __name__ == '__main__':
    main(sys.argv[1])
