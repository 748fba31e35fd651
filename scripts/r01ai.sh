set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 > gpurun_out/r01ai_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup"
KMN_PIPELINE=0 $B > gpurun_out/r01ai_serial.json 2> gpurun_out/r01ai_serial.err
$B > gpurun_out/r01ai_pipe.json 2> gpurun_out/r01ai_pipe.err
