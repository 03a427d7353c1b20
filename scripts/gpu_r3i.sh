mkdir -p gpurun_out
timeout 180 python scripts/variant_probe.py 2>&1 | grep -v Warning | tee gpurun_out/r3l_probe.txt
timeout 180 python scripts/timeline.py > gpurun_out/r3l_timeline.txt 2>&1; grep -A2 "back/epi" gpurun_out/r3l_timeline.txt | cut -c1-1100
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "bf16 or variants or select or render" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-comparator > gpurun_out/r3l_bench.json 2> gpurun_out/r3l_bench.err; tail -2 gpurun_out/r3l_bench.err; cut -c1-200 gpurun_out/r3l_bench.json
