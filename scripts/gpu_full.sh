#!/bin/bash
TAG=${1:-r2x}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 | cut -c1-600
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench_n1.json'));print(d['value'],d['e2e']['value'],d['roofline']['phase_ms_per_step'], d['gpu_launches'], d['roofline']['frac'], d.get('parity_vs_reference_cuda'), d.get('gpu_comparator',{}).get('value'))"; tail -2 gpurun_out/${TAG}_bench_n1.err
