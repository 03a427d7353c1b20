# Round-1 measurement bundle (1 GPU): tests, bench (both arms), ncu launch list + full captures.
set -x
mkdir -p gpurun_out
TAG=${1:-r1c}
python -m pytest tests -m gpu -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; cat gpurun_out/bench_${TAG}_ref.json | cut -c1-300
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -3 gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}.json
export SNB_BENCH_MIN_WARMUP=1
ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 500 --kill 1 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
ncu --set full --clock-control none --import-source on -k regex:k_back -s 40 -c 1 --kill 1 -o gpurun_out/prof_back_${TAG} python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_back.log 2>&1; tail -2 gpurun_out/ncu_back.log
ncu --set full --clock-control none --import-source on -k regex:k_front -s 40 -c 1 --kill 1 -o gpurun_out/prof_front_${TAG} python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_front.log 2>&1; tail -2 gpurun_out/ncu_front.log
ls -la gpurun_out | head -30
