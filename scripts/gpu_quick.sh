mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for cfg in "2 8 x" "2 8 1" "3 8 x" "2 12 x" "2 4 x"; do set -- $cfg
if [ "$3" = "1" ]; then export SNB_BACK_PART=1; else unset SNB_BACK_PART; fi
SNB_PIPE_DEPTH=$1 SNB_ROUTE_SMS=$2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err; python -c "
import json;d=json.load(open('gpurun_out/bench_quick.json'));print('depth $1 route_sms $2 backpart $3: ms/step',round(d['ms_per_step'],3),'value',round(d['value']/1e6,1),d['roofline']['phase_ms_per_step'])"
done
