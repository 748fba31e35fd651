set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 8 > gpurun_out/r01aj_pytest.log
python bench.py > gpurun_out/r01aj_bench.json 2> gpurun_out/r01aj_bench.err
