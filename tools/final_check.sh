set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --pipeline staged --steps 50 > gpurun_out/bench_staged_final.json 2> gpurun_out/bs.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2>gpurun_out/br.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_final_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tx_map|jakes_coef|channel_rx|crs_ls|mrc_kernel" -c 5 -f -o gpurun_out/r01_final python tools/stage_bench.py --fused --reps 1 > gpurun_out/ncu_f.log 2>&1
cat gpurun_out/pytest_final.log; cut -c1-200 gpurun_out/bench_final.json; tail -3 gpurun_out/ncu_f.log
