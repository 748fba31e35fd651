set -x
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r01s_gpu.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r01s_pytest.log
python bench.py > gpurun_out/r01s_bench.json 2> gpurun_out/r01s_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01s_ref.json 2> gpurun_out/r01s_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01s_launches.csv python bench.py --reads 20000000 --genome 50000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r01s_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_insert_staged|k_kmer_scatter|k_weight_mask" -c 3 -o gpurun_out/r01s_full python bench.py --reads 20000000 --genome 50000000 --steps 1 --warmup 0 --no-cpu --no-e2e --pipe-batches 2 > gpurun_out/r01s_full.log 2>&1
ls -la gpurun_out
cp MEASURED_PEAKS.json gpurun_out/ 2>/dev/null
