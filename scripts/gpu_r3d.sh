mkdir -p gpurun_out
for v in "" nocopy samesrc; do
  if [ -z "$v" ]; then timeout 180 python scripts/variant_probe.py; else SNB_LIB=switch_nerf_b200/variants/libsnb_$v.so timeout 180 python scripts/variant_probe.py; fi
done 2>&1 | grep -v Warning | tee gpurun_out/r3d_variants.txt
SNB_LIB=switch_nerf_b200/variants/libsnb_nocopy.so timeout 180 python scripts/timeline.py > gpurun_out/r3d_timeline_nocopy.txt 2>&1; tail -5 gpurun_out/r3d_timeline_nocopy.txt | cut -c1-300
