#!/bin/bash
# round-end validation: all GPU tests, smoke, both bench workloads, launch list + one full ncu capture of launch #2
TAG=${1:-r3p}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-600
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench_n1.json'));print(d['value'],d['e2e']['value'],d['roofline']['phase_ms_per_step'], d['gpu_launches'], d['roofline']['frac'], d.get('parity_vs_reference_cuda',{}).get('psnr_db'), d.get('gpu_comparator',{}).get('value'), d.get('cpu_baseline',{}).get('value'))"; tail -2 gpurun_out/${TAG}_bench_n1.err
timeout 600 python bench.py --workload mission_bay --steps 5 --warmup 3 --no-gpu-comparator > gpurun_out/${TAG}_bench_mission_bay.json 2> gpurun_out/${TAG}_bench_mission_bay.err; python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench_mission_bay.json'));print('mb', d['value'],d['e2e']['value'],d['roofline']['phase_ms_per_step'], d['roofline']['frac'])"; tail -2 gpurun_out/${TAG}_bench_mission_bay.err
bash scripts/gpu_launchlist.sh ${TAG}
SNB_BENCH_MIN_WARMUP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_back_ts -s 20 -c 1 --kill 1 -o gpurun_out/${TAG}_prof_back python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-comparator > gpurun_out/${TAG}_ncu_back.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_back.log
ncu -i gpurun_out/${TAG}_prof_back.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_k_back_raw.csv 2>/dev/null; ncu -i gpurun_out/${TAG}_prof_back.ncu-rep --page details > gpurun_out/${TAG}_ncu_full_k_back_details.txt 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/${TAG}_ncu_full_k_back_raw.csv')))
H=rows[0]
def col(n):
    return rows[2][H.index(n)] if n in H else None
for n in ("dram__bytes_read.sum","dram__bytes_write.sum","gpu__time_duration.sum","sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_tmem.sum"):
    print(n, col(n), rows[1][H.index(n)] if n in H else None)
PY
rm -f gpurun_out/${TAG}_prof_back.ncu-rep
