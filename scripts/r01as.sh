python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks > gpurun_out/r01as_bench.json 2> gpurun_out/r01as_bench.err
