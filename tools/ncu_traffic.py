#!/usr/bin/env python
"""profiles/rNN_traffic.json from an `ncu --set full` report of tools/stage_bench.py: per stage of the bench, the
dram__bytes_read.sum + dram__bytes_write.sum and duration of ONE launch of its dominant kernel.  bench.py reads
`roofline.traffic` from this file and refuses to run if the kernel it reports is not in it.
usage: python tools/ncu_traffic.py [--update] out.json subframes_per_launch report.ncu-rep [more.ncu-rep ...]"""
import csv
import json
import subprocess
import sys

# bench.py stage name -> substring of the CUDA kernel name
STAGE_KERNELS = {
    'tx_spectral': 'tx_spectral_kernel', 'channel_spectral': 'channel_spectral_kernel',
    'crs_ls_compact': 'crs_ls_compact_kernel', 'mrc_demap_count_compact': 'mrc_compact_kernel',
    'tx_map_ifft': 'tx_map_ifft_kernel', 'channel_rx_fft': 'channel_rx_fft_kernel',
    'crs_ls_interp': 'crs_ls_interp_kernel', 'mrc_demap_count': 'mrc_kernel', 'channel_tdl': 'tdl_kernel',
    'rx_fft': 'rx_fft_kernel',
}


def to_bytes(v, unit):
    return float(v) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]


def main():
    # --update: keep the entries of an existing out.json for kernels the given reports do not hold
    args = [a for a in sys.argv[1:] if a != '--update']
    out, per_launch, reps = args[0], int(args[1]), args[2:]
    kernels, old_reports = {}, []
    if '--update' in sys.argv[1:]:
        prev = json.load(open(out))
        kernels = prev['kernels']
        old_reports = sorted({k['report'] for k in kernels.values()})
    for rep in reps:
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[col['Kernel Name']]
            for stage, sub in STAGE_KERNELS.items():
                if sub in name and (stage != 'rx_fft' or 'channel' not in name) and (stage != 'mrc_demap_count' or 'compact' not in name):
                    rd = to_bytes(r[col['dram__bytes_read.sum']], units[col['dram__bytes_read.sum']])
                    wr = to_bytes(r[col['dram__bytes_write.sum']], units[col['dram__bytes_write.sum']])
                    dur = float(r[col['gpu__time_duration.sum']])
                    du = units[col['gpu__time_duration.sum']]
                    dur_us = dur * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'msecond': 1e3, 'nsecond': 1e-3}.get(du, 1.0)
                    kernels[stage] = {'kernel': name[:120], 'report': rep.split('/')[-1], 'dram_bytes': rd + wr, 'dram_read': rd,
                                      'dram_write': wr, 'duration_us': dur_us}
    json.dump({'source': 'ncu --set full --clock-control none of tools/stage_bench.py --pipeline spectral|fused|staged, one launch '
                         'per kernel (' + ', '.join(sorted(set(old_reports) | {r.split('/')[-1] for r in reps})) + ')',
               'subframes_per_launch': per_launch, 'kernels': kernels}, open(out, 'w'), indent=1)
    print(json.dumps({k: round(v['dram_bytes'] / 1e9, 3) for k, v in kernels.items()}))


if __name__ == '__main__':
    main()
