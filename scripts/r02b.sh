set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02b_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup"
KMN_PIPELINE=0 $B > gpurun_out/r02b_def_serial.json 2> gpurun_out/r02b_def_serial.err
KMN_PIPELINE=0 KMN_SCATTER_TPB=1024 KMN_SCATTER_CTAS=1 $B --no-checks > gpurun_out/r02b_t1024_serial.json 2> gpurun_out/r02b_t1024_serial.err
KMN_PIPELINE=0 KMN_SCATTER_TPB=256 KMN_SCATTER_CTAS=4 $B --no-checks > gpurun_out/r02b_t256_serial.json 2> gpurun_out/r02b_t256_serial.err
KMN_PIPELINE=0 KMN_RING=0 $B --no-checks > gpurun_out/r02b_noring_serial.json 2> gpurun_out/r02b_noring_serial.err
KMN_PIPELINE=0 $B --no-checks --slice-mb 32 > gpurun_out/r02b_s32_serial.json 2> gpurun_out/r02b_s32_serial.err
KMN_PIPELINE=0 $B --no-checks --slice-mb 16 > gpurun_out/r02b_s16_serial.json 2> gpurun_out/r02b_s16_serial.err
$B --no-checks > gpurun_out/r02b_def_pipe.json 2> gpurun_out/r02b_def_pipe.err
