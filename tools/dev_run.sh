python tools/time_spectral.py 4096 2>&1 | tail -1
python -m pytest tests/test_gpu_spectral.py -x -q -k "Pedestrian_A-4-14 or counts_exactly" 2>&1 | tail -3
