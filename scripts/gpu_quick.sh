mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err; python -c "
import json;d=json.load(open('gpurun_out/bench_quick.json'));print('OVERLAP   ms/step',d['ms_per_step'],'value',d['value'],d['roofline']['phase_ms_per_step'],'launches',d['gpu_launches'], 'frac', d['roofline']['frac'])"
python scripts/timeline.py 2>&1 | grep -A3 "front/epi\|back/epi" | cut -c1-700
