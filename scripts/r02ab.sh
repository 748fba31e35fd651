set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r02ab_pytest.log
timeout 900 python bench.py > gpurun_out/r02ab_bench.json 2> gpurun_out/r02ab_bench.err
tail -c 3000 gpurun_out/r02ab_bench.err > gpurun_out/r02ab_bench.err.tail; rm -f gpurun_out/r02ab_bench.err
