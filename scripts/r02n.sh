set -x
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r02n_pytest.log
B="timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks"
$B > gpurun_out/r02n_def.json 2> gpurun_out/r02n_def.err
$B --pipe-batches 2 > gpurun_out/r02n_pb2.json 2> gpurun_out/r02n_pb2.err
$B --pipe-batches 3 > gpurun_out/r02n_pb3.json 2> gpurun_out/r02n_pb3.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv > gpurun_out/r02n_mem.txt
for f in gpurun_out/r02n_*.err; do tail -c 4000 $f > $f.tail; rm -f $f; done
