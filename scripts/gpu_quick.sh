mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for cfg in "2 8" "2 0" "1 0" "3 8" "2 16" "1 8"; do set -- $cfg
SNB_PIPE_DEPTH=$1 SNB_ROUTE_SMS=$2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err; python -c "
import json;d=json.load(open('gpurun_out/bench_quick.json'));print('depth $1 route_sms $2: ms/step',round(d['ms_per_step'],3),'value',round(d['value']/1e6,1),d['roofline']['phase_ms_per_step'])"
done
