#!/bin/bash
# bf16 contract tests against the reference's CUDA-autocast goldens + smoke + bench line
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bf16 or perturb" -s 2>&1 | tail -30 > gpurun_out/${TAG}_pytest_bf16.log; tail -12 gpurun_out/${TAG}_pytest_bf16.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | cut -c1-700
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; cut -c1-1500 gpurun_out/${TAG}_bench_n1.json; tail -3 gpurun_out/${TAG}_bench_n1.err
