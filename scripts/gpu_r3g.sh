mkdir -p gpurun_out
timeout 180 python scripts/variant_probe.py 2>&1 | grep -v Warning | tee gpurun_out/r3g_probe.txt
timeout 180 python scripts/timeline.py > gpurun_out/r3g_timeline.txt 2>&1; grep -A2 "back/mma" gpurun_out/r3g_timeline.txt | cut -c1-900
timeout 400 python -m pytest tests/test_gpu_parity.py -q -x -k "bf16 or variants or select" 2>&1 | tail -5
