#!/bin/bash
# 1-GPU bundle for the TMEM-operand (TS) variant of launch #2: layout probe, parity, A/B bench.
TAG=${1:-r1h}
mkdir -p gpurun_out
timeout 120 python scripts/ts_probe.py > gpurun_out/${TAG}_ts_probe.json 2> gpurun_out/${TAG}_ts_probe.err; cat gpurun_out/${TAG}_ts_probe.json; tail -3 gpurun_out/${TAG}_ts_probe.err
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "umma_selftest or tmem_operand or cta_pair or full_chunk" > gpurun_out/${TAG}_ts_tests.log 2>&1
tail -25 gpurun_out/${TAG}_ts_tests.log
SNB_TS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_ts.json 2> gpurun_out/${TAG}_bench_ts.err
tail -c 900 gpurun_out/${TAG}_bench_ts.json; tail -5 gpurun_out/${TAG}_bench_ts.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
tail -c 900 gpurun_out/${TAG}_bench_default.json; tail -5 gpurun_out/${TAG}_bench_default.err
