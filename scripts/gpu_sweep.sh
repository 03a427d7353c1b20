#!/bin/bash
# quick A/B of the launch-partition knobs (bench.py, 10 steps each)
TAG=${1:-r1m}
mkdir -p gpurun_out
for v in "SNB_ROUTE_SMS=12" "SNB_ROUTE_SMS=28" "SNB_ROUTE_SMS=36" "SNB_PIPE_DEPTH=3" "SNB_BACK_PART=1"; do
  name=$(echo $v | tr '= ' '__')
  env $v timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${name}.json").read().strip().splitlines()[-1])
    print("${v}", round(d["value"]/1e6,1), "M/s", round(d["ms_per_step"],3), "ms", d["roofline"]["phase_ms_per_step"], round(d["roofline"]["frac"],3))
except Exception as e:
    print("${v}", "FAILED", e); print(open("gpurun_out/${TAG}_bench_${name}.err").read()[-800:])
PY
done
