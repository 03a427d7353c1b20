#!/bin/bash
# 2-GPU bundle: expert-parallel parity, DP vs EP bench, NCCL a2a floor, unfused-PyTorch comparator.
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_ep2.sh r1f'
TAG=${1:-r1f}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
timeout 420 python -m pytest tests/test_gpu_parity.py -x -q -k expert_parallel > gpurun_out/${TAG}_ep_test.log 2>&1
echo "ep test exit $?" >> gpurun_out/${TAG}_ep_test.log
tail -30 gpurun_out/${TAG}_ep_test.log
timeout 240 $TR --master-port 29751 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n2_dp.json 2> gpurun_out/${TAG}_bench_n2_dp.err
tail -c 600 gpurun_out/${TAG}_bench_n2_dp.json
timeout 240 $TR --master-port 29752 bench.py --gpus 2 --steps 10 --warmup 3 --parallelism ep > gpurun_out/${TAG}_bench_n2_ep.json 2> gpurun_out/${TAG}_bench_n2_ep.err
tail -c 600 gpurun_out/${TAG}_bench_n2_ep.json; tail -5 gpurun_out/${TAG}_bench_n2_ep.err
timeout 120 $TR --master-port 29753 scripts/nccl_a2a_baseline.py > gpurun_out/${TAG}_nccl_a2a_n2.json 2> gpurun_out/${TAG}_nccl_a2a_n2.err
cat gpurun_out/${TAG}_nccl_a2a_n2.json
CUDA_VISIBLE_DEVICES=0 timeout 240 python scripts/torch_unfused_gpu.py --steps 5 --warmup 2 > gpurun_out/${TAG}_torch_unfused.json 2> gpurun_out/${TAG}_torch_unfused.err
cat gpurun_out/${TAG}_torch_unfused.json; tail -3 gpurun_out/${TAG}_torch_unfused.err
