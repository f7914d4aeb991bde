#!/bin/bash
# Round-end ncu pass on one GPU box: `--set full` capture of the spectral pipeline's stage kernels (after the same command
# exited 0 without the profiler), its summary and traffic table, then the launch list of the default bench command.
tag=${1:-r02}
bash tools/ncu_capture.sh spectral $tag || exit 1
python tools/ncu_summary.py gpurun_out/${tag}_spectral.ncu-rep > gpurun_out/${tag}_spectral_ncu_summary.md
cp profiles/${tag}_traffic.json gpurun_out/${tag}_traffic.json
python tools/ncu_traffic.py --update gpurun_out/${tag}_traffic.json 4096 gpurun_out/${tag}_spectral.ncu-rep
python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/plain_ll.log 2>&1 || { echo "plain bench failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_spectral_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/ncu_ll.log 2>&1
python tools/launch_share.py gpurun_out/${tag}_spectral_launches.csv 2>&1 | tail -12
