mkdir -p gpurun_out
SNB_LIB=switch_nerf_b200/variants/libsnb_tlfine.so timeout 180 python scripts/timeline.py > gpurun_out/r3e_timeline_fine_nocopy.txt 2>&1; grep -A3 "back/mma" gpurun_out/r3e_timeline_fine_nocopy.txt | cut -c1-1500
