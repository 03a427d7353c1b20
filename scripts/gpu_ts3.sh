#!/bin/bash
TAG=${1:-r1l}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; tail -6 gpurun_out/${TAG}_gpu_tests.log
for v in default "SNB_TS_FRONT=0"; do
  name=$(echo $v | tr '= ' '__')
  env $( [ "$v" = default ] || echo $v ) timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${name}.json").read().strip().splitlines()[-1])
    print("${v}", round(d["value"]/1e6,1), "M/s", round(d["ms_per_step"],3), "ms", d["roofline"]["phase_ms_per_step"], round(d["roofline"]["frac"],3))
except Exception as e:
    print("${v}", "FAILED", e); print(open("gpurun_out/${TAG}_bench_${name}.err").read()[-1500:])
PY
done
timeout 200 python scripts/timeline.py > gpurun_out/${TAG}_timeline.txt 2>&1
grep -A2 "back/epi\|front/epi" gpurun_out/${TAG}_timeline.txt | cut -c1-1000
