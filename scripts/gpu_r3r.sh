mkdir -p gpurun_out
for cfg in "" "SNB_PIPE_DEPTH=3" "SNB_ROUTE_SMS=4" "SNB_ROUTE_SMS=12" "SNB_ROUTE_SMS=16 SNB_PIPE_DEPTH=3" "SNB_BACK_PART=1"; do
  env $cfg timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-comparator 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', round(d['value']/1e6,1), d['ms_per_step'], d['roofline']['phase_ms_per_step'])"
done | tee gpurun_out/r3r_pipe_knobs.txt
