set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r01u_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
$B > gpurun_out/r01u_def.json 2> gpurun_out/r01u_def.err
KMN_PIPELINE=0 $B > gpurun_out/r01u_def_serial.json 2> gpurun_out/r01u_def_serial.err
KMN_LIB_VARIANT=c3 KMN_PIPELINE=0 $B > gpurun_out/r01u_c3_serial.json 2> gpurun_out/r01u_c3_serial.err
KMN_LIB_VARIANT=u2 KMN_PIPELINE=0 $B > gpurun_out/r01u_u2_serial.json 2> gpurun_out/r01u_u2_serial.err
KMN_LIB_VARIANT=u8c2 KMN_PIPELINE=0 $B > gpurun_out/r01u_u8c2_serial.json 2> gpurun_out/r01u_u8c2_serial.err
KMN_PIPELINE=0 $B --slice-mb 128 > gpurun_out/r01u_s128_serial.json 2> gpurun_out/r01u_s128_serial.err
KMN_PIPELINE=0 $B --slice-mb 32 > gpurun_out/r01u_s32_serial.json 2> gpurun_out/r01u_s32_serial.err
KMN_PIPELINE=0 KMN_NO_L2_HINTS=1 $B > gpurun_out/r01u_nohint_serial.json 2> gpurun_out/r01u_nohint_serial.err
