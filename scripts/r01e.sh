set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r01e_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
$B > gpurun_out/r01e_m1.json 2> gpurun_out/r01e_m1.err
$B --slice-mb 16 > gpurun_out/r01e_m2.json 2> gpurun_out/r01e_m2.err
$B --pipe-batches 2 > gpurun_out/r01e_m3.json 2> gpurun_out/r01e_m3.err
ncu --set full --clock-control none --import-source on -k regex:"k_kmer_scatter|k_weight_mask" -c 2 -f -o gpurun_out/r01e_p1_full python bench.py --reads 20000000 --genome 50000000 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r01e_ncu.log 2>&1
