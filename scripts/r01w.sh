set -x
export KMN_LIB_VARIANT=wm
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r01w_pytest.log
TR="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu --no-e2e"
$TR > gpurun_out/r01w_p2p.json 2> gpurun_out/r01w_p2p.err
KMN_P2P=0 $TR > gpurun_out/r01w_nccl.json 2> gpurun_out/r01w_nccl.err
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r01w_1gpu.json 2> gpurun_out/r01w_1gpu.err
nvidia-smi topo -m > gpurun_out/r01w_topo.txt 2>&1
