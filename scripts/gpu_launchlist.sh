#!/bin/bash
TAG=${1:-r2r}
mkdir -p gpurun_out
SNB_BENCH_MIN_WARMUP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_bench_1step.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-comparator > /dev/null 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open('gpurun_out/${TAG}_launches_bench_1step.csv')))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
H=rows[hdr]; ki=H.index('Kernel Name'); vi=H.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[hdr+1:]:
    if len(r)<=vi: continue
    n=r[ki].split('(')[0][:50]
    a=agg.setdefault(n,[0,0.0,[]]); a[0]+=1; a[1]+=float(r[vi].replace(',','')); a[2].append(float(r[vi].replace(',','')))
for n,(c,t,l) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:8]:
    l.sort()
    print(f"  {n:40s} {c:5d} avg {t/c/1e3:8.1f} us  median {l[len(l)//2]/1e3:8.1f} min {l[0]/1e3:.1f} max {l[-1]/1e3:.1f}")
PY
