set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02k_pytest.log
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --workload c5b > gpurun_out/r02k_c5b.json 2> gpurun_out/r02k_c5b.err
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --workload c4 > gpurun_out/r02k_c4.json 2> gpurun_out/r02k_c4.err
timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu --no-lookup --no-checks > gpurun_out/r02k_e2e.json 2> gpurun_out/r02k_e2e.err
for f in gpurun_out/r02k_*.err; do tail -c 6000 $f > $f.tail; rm -f $f; done
