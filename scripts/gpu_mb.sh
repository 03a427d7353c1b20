#!/bin/bash
TAG=${1:-r2w}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-400
timeout 600 python bench.py --workload mission_bay --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_mission_bay.json 2> gpurun_out/${TAG}_bench_mission_bay.err; cut -c1-2500 gpurun_out/${TAG}_bench_mission_bay.json; tail -3 gpurun_out/${TAG}_bench_mission_bay.err
