#!/usr/bin/env python
"""Per-source-line instruction counts / stall samples for one kernel of an .ncu-rep.
usage: tools/ncu_lines.py rep kernel_substring [topN]"""
import csv, io, subprocess, sys, collections
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
text = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                      capture_output=True, text=True).stdout
cur_file = None; in_kernel = False; cols = None
agg = collections.OrderedDict()
for row in csv.reader(io.StringIO(text)):
    if not row: continue
    if row[0] == 'File Path': cur_file = row[1]; continue
    if row[0] == 'Function Name': in_kernel = pat in row[1]; continue
    if row[0] == 'Kernel Name': in_kernel = False; continue
    if row[0] == 'Line No': cols = {h: i for i, h in enumerate(row)}; continue
    if not in_kernel or cols is None: continue
    if row[0] and row[0].isdigit():
        key = (cur_file.split('/')[-1], int(row[0]))
        try:
            n = int(row[cols['Instructions Executed']] or 0); s = int(row[cols['# Samples']] or 0)
        except (ValueError, KeyError):
            continue
        src = row[1].strip()[:90]
        a = agg.setdefault(key, [0, 0, src]); a[0] += n; a[1] += s
tot = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
print(f'total warp instr {tot}, samples {ts}')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{k[0]:12s}:{k[1]:4d} {100*v[0]/tot:5.1f}% instr {100*v[1]/ts:5.1f}% stall | {v[2]}')
