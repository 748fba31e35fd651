set -x
timeout 800 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "count_table_synthetic or many_small or trim_scores or saturation or meraculous" > gpurun_out/r01am_memcheck.log 2>&1
tail -n 30 gpurun_out/r01am_memcheck.log
