mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json
python scripts/timeline.py 2>&1 | cut -c1-900
