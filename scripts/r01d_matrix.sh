set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r01d_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
KMN_NO_PIPELINE=1 KMN_PARSE_TPB=512 KMN_INSERT_CTAS=8 $B > gpurun_out/r01d_m1.json 2> gpurun_out/r01d_m1.err
KMN_PARSE_TPB=384 KMN_INSERT_CTAS=1 $B > gpurun_out/r01d_m2.json 2> gpurun_out/r01d_m2.err
KMN_PARSE_TPB=256 KMN_INSERT_CTAS=2 $B > gpurun_out/r01d_m3.json 2> gpurun_out/r01d_m3.err
KMN_PARSE_TPB=512 KMN_INSERT_CTAS=8 $B > gpurun_out/r01d_m4.json 2> gpurun_out/r01d_m4.err
KMN_NO_PIPELINE=1 KMN_PARSE_TPB=384 KMN_INSERT_CTAS=1 $B > gpurun_out/r01d_m5.json 2> gpurun_out/r01d_m5.err
KMN_PARSE_TPB=384 KMN_INSERT_CTAS=1 $B --pipe-batches 16 > gpurun_out/r01d_m6.json 2> gpurun_out/r01d_m6.err
KMN_PARSE_TPB=384 KMN_INSERT_CTAS=1 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r01d_m7.json 2> gpurun_out/r01d_m7.err
