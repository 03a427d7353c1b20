mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err; python -c "
import json;d=json.load(open('gpurun_out/bench_quick.json'));print('OVERLAP   ms/step',d['ms_per_step'],'value',d['value'],d['roofline']['phase_ms_per_step'],'launches',d['gpu_launches'], 'frac', d['roofline']['frac'])"
SNB_NO_OVERLAP=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick2.json 2> gpurun_out/bench_quick2.err; tail -3 gpurun_out/bench_quick2.err; python -c "
import json;d=json.load(open('gpurun_out/bench_quick2.json'));print('SEQUENTIAL ms/step',d['ms_per_step'],'value',d['value'],d['roofline']['phase_ms_per_step'])"
