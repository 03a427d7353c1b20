#!/bin/bash
# select routing: parity tests, bench, launch list
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/${TAG}_pytest_gpu.log; tail -12 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | cut -c1-600
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; cut -c1-300 gpurun_out/${TAG}_bench_n1.json; tail -3 gpurun_out/${TAG}_bench_n1.err
for sms in 8; do
SNB_ROUTE_SMS=$sms timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-comparator > gpurun_out/${TAG}_bench_route_sms_$sms.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench_route_sms_$sms.json'));print('route_sms',$sms,d['value'],d['roofline']['phase_ms_per_step'])"
done
SNB_BENCH_MIN_WARMUP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_bench_1step.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-comparator > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/'+'${TAG}'+'_launches_bench_1step.csv')) if len(r)>5]
PY
grep -c . gpurun_out/${TAG}_launches_bench_1step.csv
