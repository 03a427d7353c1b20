#!/bin/bash
TAG=${1:-r2e}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "route" 2>&1 | tail -5
timeout 300 python scripts/timeline.py > gpurun_out/${TAG}_timeline.txt 2>&1; grep "k_select" gpurun_out/${TAG}_timeline.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-comparator > gpurun_out/${TAG}_bench_n1.json 2>/dev/null; cut -c1-200 gpurun_out/${TAG}_bench_n1.json
