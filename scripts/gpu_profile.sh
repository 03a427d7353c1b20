mkdir -p gpurun_out
export SNB_BENCH_MIN_WARMUP=1
TAG=${1:-r1b}
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 700 --kill 1 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
ncu --set full --clock-control none --import-source on -k regex:k_back -s 40 -c 1 --kill 1 -o gpurun_out/prof_back_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_back.log 2>&1; tail -2 gpurun_out/ncu_back.log
ncu --set full --clock-control none --import-source on -k regex:k_front -s 40 -c 1 --kill 1 -o gpurun_out/prof_front_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_front.log 2>&1; tail -2 gpurun_out/ncu_front.log
ls -la gpurun_out | head -20
