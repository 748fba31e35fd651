set -x
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02f_pytest.log
P="python bench.py --reads 25000000 --genome 62500000 --steps 1 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks"
$P > gpurun_out/r02f_small.json 2> gpurun_out/r02f_small.err
ncu --set full --clock-control none --import-source on -k regex:k_count_slices -s 4 -c 1 -f -o gpurun_out/r02f_count $P > gpurun_out/r02f_ncu_count.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_slice_split -s 4 -c 1 -f -o gpurun_out/r02f_split $P > gpurun_out/r02f_ncu_split.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_kmer_scatter -s 4 -c 1 -f -o gpurun_out/r02f_scatter $P > gpurun_out/r02f_ncu_scatter.log 2>&1
