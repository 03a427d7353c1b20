mkdir -p gpurun_out
for v in "" nogroup nst4 nst6diag accspin; do
  if [ -z "$v" ]; then timeout 180 python scripts/variant_probe.py; else SNB_LIB=switch_nerf_b200/variants/libsnb_$v.so timeout 180 python scripts/variant_probe.py; fi
done 2>&1 | grep -v Warning | tee gpurun_out/r3c_variants.txt
timeout 180 python scripts/timeline.py > gpurun_out/r3c_timeline.txt 2>&1; tail -5 gpurun_out/r3c_timeline.txt | cut -c1-400
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "bf16" 2>&1 | tail -5
