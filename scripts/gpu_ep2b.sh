#!/bin/bash
# 2-GPU: expert-parallel parity + DP / EP bench with the current kernels
TAG=${1:-r1q}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 420 python -m pytest tests/test_gpu_parity.py -x -q -k expert_parallel > gpurun_out/${TAG}_ep_test.log 2>&1
tail -4 gpurun_out/${TAG}_ep_test.log
timeout 240 $TR --master-port 29751 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n2_dp.json 2> gpurun_out/${TAG}_bench_n2_dp.err
timeout 240 $TR --master-port 29752 bench.py --gpus 2 --steps 10 --warmup 3 --parallelism ep > gpurun_out/${TAG}_bench_n2_ep.json 2> gpurun_out/${TAG}_bench_n2_ep.err
python - <<PY
import json
for n in ("dp","ep"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_n2_%s.json" % n).read().strip().splitlines()[-1])
        print(n, round(d["value"]/1e6,1), "M/s", round(d["ms_per_step"],3), "ms e2e", round(d["e2e"]["value"]/1e6,1), d["roofline"]["phase_ms_per_step"])
    except Exception as e:
        print(n, "FAILED", e); print(open("gpurun_out/${TAG}_bench_n2_%s.err" % n).read()[-1500:])
PY
