set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err; tail -3 gpurun_out/bench_r1_a.err; cat gpurun_out/bench_r1_a.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref.json 2>&1; cat gpurun_out/bench_r1_ref.json
export SNB_BENCH_MIN_WARMUP=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --kill 1 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
ncu --set full --clock-control none --import-source on -k regex:k_back -s 3 -c 2 --kill 1 -o gpurun_out/prof_back_r1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_back.log 2>&1; tail -2 gpurun_out/ncu_back.log
ncu --set full --clock-control none --import-source on -k regex:k_front -s 3 -c 1 --kill 1 -o gpurun_out/prof_front_r1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_front.log 2>&1; tail -2 gpurun_out/ncu_front.log
ls -la gpurun_out
