mkdir -p gpurun_out
for v in "" ns10 accspin; do
  if [ -z "$v" ]; then timeout 180 python scripts/variant_probe.py; else SNB_LIB=switch_nerf_b200/variants/libsnb_$v.so timeout 180 python scripts/variant_probe.py; fi
done 2>&1 | grep -v Warning | tee gpurun_out/r3t_backoff_variants.txt
