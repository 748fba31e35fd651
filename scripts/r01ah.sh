set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 > gpurun_out/r01ah_pytest.log
python bench.py > gpurun_out/r01ah_bench.json 2> gpurun_out/r01ah_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01ah_ref.json 2> gpurun_out/r01ah_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01ah_launches.csv python bench.py --reads 20000000 --genome 50000000 --steps 1 --warmup 1 --no-cpu --no-e2e --no-lookup > gpurun_out/r01ah_launches.log 2>&1
KMN_PIPELINE=0 ncu --set full --clock-control none --import-source on -k regex:k_insert_staged -s 1 -c 1 -o gpurun_out/r01ah_insert python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-lookup > gpurun_out/r01ah_ncu.log 2>&1
python __graft_entry__.py smoke > gpurun_out/r01ah_smoke.log 2>&1
