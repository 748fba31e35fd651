set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02ai_pytest.log
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --workload c5a > gpurun_out/r02ai_c5a.json 2> gpurun_out/r02ai_c5a.err
tail -c 2000 gpurun_out/r02ai_c5a.err > gpurun_out/r02ai_c5a.err.tail; rm -f gpurun_out/r02ai_c5a.err
