#!/bin/bash
TAG=${1:-r2k}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "route or bf16" 2>&1 | tail -3
python scripts/select_probe.py
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-comparator > gpurun_out/${TAG}_bench_n1.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench_n1.json'));print(d['value'],d['roofline']['phase_ms_per_step'], d['gpu_launches'])"
