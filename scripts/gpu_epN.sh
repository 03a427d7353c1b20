#!/bin/bash
# expert-parallel parity (tests/ep_worker.py) + bench with the ep leg at N GPUs
N=${1:-2}; TAG=${2:-r2m}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 tests/ep_worker.py > gpurun_out/${TAG}_ep_parity_n$N.log 2>&1; tail -6 gpurun_out/${TAG}_ep_parity_n$N.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus $N --steps 10 --warmup 3 $BENCH_EXTRA > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${TAG}_bench_n$N.json') if l.startswith('{')][-1])
    print('dp', d['value'], d['ms_per_step'], 'ep', d.get('ep'))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/${TAG}_bench_n$N.err').read()[-2000:])
PY
