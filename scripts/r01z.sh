set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r01z_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
KMN_PIPELINE=0 $B > gpurun_out/r01z_serial.json 2> gpurun_out/r01z_serial.err
$B --no-lookup > gpurun_out/r01z_pipe.json 2> gpurun_out/r01z_pipe.err
KMN_PIPELINE=0 $B --slice-mb 128 --no-lookup > gpurun_out/r01z_s128_serial.json 2> gpurun_out/r01z_s128_serial.err
KMN_PIPELINE=0 ncu --set full --clock-control none --import-source on -k regex:k_weight_mask -s 1 -c 1 -o gpurun_out/r01z_weight python bench.py --reads 20000000 --genome 50000000 --steps 1 --warmup 0 --no-cpu --no-e2e --no-lookup > gpurun_out/r01z_ncu.log 2>&1
