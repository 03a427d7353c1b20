mkdir -p gpurun_out
timeout 180 python scripts/variant_probe.py 2>&1 | grep -v Warning | tee gpurun_out/r3h_probe.txt
timeout 180 python scripts/timeline.py > gpurun_out/r3h_timeline.txt 2>&1; grep -A2 "back/mma" gpurun_out/r3h_timeline.txt | cut -c1-700
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "bf16 or variants or select or mip or wide" 2>&1 | tail -4
timeout 300 python bench.py --workload mission_bay --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-comparator > gpurun_out/r3h_bench_mb.json 2> gpurun_out/r3h_bench_mb.err; tail -2 gpurun_out/r3h_bench_mb.err; cut -c1-600 gpurun_out/r3h_bench_mb.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-comparator > gpurun_out/r3h_bench.json 2> gpurun_out/r3h_bench.err; tail -2 gpurun_out/r3h_bench.err; cut -c1-900 gpurun_out/r3h_bench.json
