mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json
python scripts/cf_sweep.py 2>&1 | tail -30
