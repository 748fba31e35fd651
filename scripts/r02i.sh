set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02i_pytest.log
B="timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks"
$B > gpurun_out/r02i_def.json 2> gpurun_out/r02i_def.err
timeout 1200 python bench.py > gpurun_out/r02i_full.json 2> gpurun_out/r02i_full.err
