#!/bin/bash
TAG=${1:-r2u}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "variants" 2>&1 | tail -8 | cut -c1-600
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "mip" 2>&1 | grep -E "AssertionError: \{|passed|failed|print|mip_" | cut -c1-900
