set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r01v_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
$B > gpurun_out/r01v_def.json 2> gpurun_out/r01v_def.err
KMN_PIPELINE=0 $B > gpurun_out/r01v_def_serial.json 2> gpurun_out/r01v_def_serial.err
for v in u4c3 u4c2 u1c5 u3c3; do
KMN_LIB_VARIANT=$v KMN_PIPELINE=0 $B > gpurun_out/r01v_${v}_serial.json 2> gpurun_out/r01v_${v}_serial.err
done
