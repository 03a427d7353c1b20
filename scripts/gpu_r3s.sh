mkdir -p gpurun_out
timeout 180 python scripts/timeline.py > gpurun_out/r3s_timeline.txt 2>&1; grep -A2 "front/epi" gpurun_out/r3s_timeline.txt | cut -c1-500
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bg.py -q -x -k "bf16 or variants or select or render or bg or mip" 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-comparator > gpurun_out/r3s_bench.json 2> gpurun_out/r3s_bench.err; tail -2 gpurun_out/r3s_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/r3s_bench.json').read().strip().splitlines()[-1]);print(d['value']/1e6,d['ms_per_step'],d['roofline']['phase_ms_per_step'],d.get('parity_vs_reference_cuda',{}).get('psnr_db'))"
timeout 300 python bench.py --workload mission_bay --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-comparator 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('mb',d['value']/1e6,d['ms_per_step'],d['roofline']['phase_ms_per_step'])"
