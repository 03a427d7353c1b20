mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -12
for v in "" "SNB_GATHER_H=1"; do
env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err; python -c "
import json;d=json.load(open('gpurun_out/bench_quick.json'));print('[$v] ms/step',round(d['ms_per_step'],3),'value',round(d['value']/1e6,1),d['roofline']['phase_ms_per_step'])"
done
