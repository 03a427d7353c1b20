#!/bin/bash
# 1-GPU: full GPU suite with the TS default, A/B bench of the launch-#2 variants, timeline of the pair variant
TAG=${1:-r1k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; tail -6 gpurun_out/${TAG}_gpu_tests.log
for v in default "SNB_CG_BACK=2" "SNB_CG=2"; do
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
SNB_CG_BACK=2 timeout 200 python scripts/timeline.py > gpurun_out/${TAG}_timeline_ts_pair.txt 2>&1
grep -A2 "back/mma" gpurun_out/${TAG}_timeline_ts_pair.txt | cut -c1-1200
