#!/bin/bash
# ncu --set full capture of one launch of every stage kernel of a pipeline (spectral | fused | staged), taken only
# after the same command has exited 0 without the profiler.  usage: bash tools/ncu_capture.sh spectral r02
pipe=${1:-spectral}; tag=${2:-r02}
case $pipe in
  spectral) rx='tx_spectral|spectral_coef|channel_spectral|mrc_compact'; n=4 ;;
  fused)    rx='tx_map_ifft|jakes_coef|channel_rx_fft|crs_ls_interp|mrc_kernel'; n=5 ;;
  staged)   rx='tx_map_ifft|jakes_coef|tdl_kernel|rx_fft_kernel|crs_ls_interp|mrc_kernel'; n=6 ;;
esac
python tools/stage_bench.py --pipeline $pipe --reps 1 > gpurun_out/plain_$pipe.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$pipe.log; exit 1; }
cat gpurun_out/plain_$pipe.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $n -f -o gpurun_out/${tag}_$pipe \
    python tools/stage_bench.py --pipeline $pipe --reps 1 > gpurun_out/ncu_$pipe.log 2>&1
tail -2 gpurun_out/ncu_$pipe.log; ls -la gpurun_out/${tag}_$pipe.ncu-rep
