mkdir -p gpurun_out
SNB_CG=2 timeout 300 python -m pytest tests -m gpu -q -x -k "bf16" 2>&1 | tail -3
SNB_CG=2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err; python -c "
import json;d=json.load(open('gpurun_out/bench_quick.json'));print('CG 2: ms/step',round(d['ms_per_step'],3),'value',round(d['value']/1e6,1),d['roofline']['phase_ms_per_step'])"
true
