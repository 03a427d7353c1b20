#!/bin/bash
# Round-2 first GPU call: goldens of the UNMODIFIED reference under cuda autocast + first comparison of the tcgen05 path
# with them, regression (tests, smoke, bench), diagnostics of the cta_group::2 variant of launch #2.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_r2a.sh r2a'
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 400 python -m oracle.make_golden_cuda --out gpurun_out/golden_cuda > gpurun_out/${TAG}_make_golden_cuda.log 2>&1; tail -3 gpurun_out/${TAG}_make_golden_cuda.log
timeout 300 python scripts/compare_cuda_golden.py gpurun_out/golden_cuda > gpurun_out/${TAG}_compare_cuda_golden.json 2> gpurun_out/${TAG}_compare_cuda_golden.err; tail -12 gpurun_out/${TAG}_compare_cuda_golden.err
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -c 600 gpurun_out/${TAG}_bench_n1.json
SNB_CG_BACK=2 timeout 200 python scripts/timeline.py > gpurun_out/${TAG}_timeline_ts_pair.txt 2>&1; tail -5 gpurun_out/${TAG}_timeline_ts_pair.txt
