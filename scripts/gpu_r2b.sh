#!/bin/bash
# goldens of the unmodified reference under cuda autocast + comparison + reference-in-the-loop tests
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 600 python -m oracle.make_golden_cuda --out gpurun_out/golden_cuda > gpurun_out/${TAG}_make_golden_cuda.log 2>&1; tail -4 gpurun_out/${TAG}_make_golden_cuda.log
timeout 300 python scripts/compare_cuda_golden.py gpurun_out/golden_cuda > gpurun_out/${TAG}_compare_cuda_golden.json 2> gpurun_out/${TAG}_compare_cuda_golden.err; tail -3 gpurun_out/${TAG}_compare_cuda_golden.err | cut -c1-600
timeout 600 python -m pytest tests/test_reference_in_loop.py -m gpu -q 2>&1 | tail -15
