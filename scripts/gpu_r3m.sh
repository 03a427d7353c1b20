mkdir -p gpurun_out
timeout 180 python scripts/timeline.py > gpurun_out/r3m_timeline.txt 2>&1; grep -A2 "back/epi" gpurun_out/r3m_timeline.txt | cut -c1-1300; grep -A2 "back/mma" gpurun_out/r3m_timeline.txt | cut -c1800-3000
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "render" 2>&1 | tail -3
