set -x
timeout 800 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "count_table_synthetic or many_small or trim_scores" > gpurun_out/r01an_racecheck.log 2>&1
timeout 800 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "count_table_synthetic or trim_scores" > gpurun_out/r01an_synccheck.log 2>&1
