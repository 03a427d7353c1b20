#!/bin/bash
# First GPU call of the next round: the diagnostics the round-1 budget did not cover.
#   gpurun --timeout 900 -- 'bash scripts/gpu_round2_first.sh r2a'
TAG=${1:-r2a}
mkdir -p gpurun_out
# 1. regression: whole GPU suite + smoke + bench (TS kernels are the default)
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -c 400 gpurun_out/${TAG}_bench_n1.json
# 2. why are cta_group::2 pairs slow inside k_back_ts although a pair MMA issues every 138 clk in isolation
#    (profiles/r1s_*)?  timeline of CTA 0 (leader) in pair mode + ncu of the pair kernel
SNB_CG_BACK=2 timeout 200 python scripts/timeline.py > gpurun_out/${TAG}_timeline_ts_pair.txt 2>&1
export SNB_BENCH_MIN_WARMUP=1
SNB_CG_BACK=2 ncu --set full --clock-control none --import-source on -k regex:k_back -s 40 -c 1 --kill 1 -o gpurun_out/prof_back_pair_${TAG} \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_back_pair.log 2>&1; tail -2 gpurun_out/ncu_back_pair.log
# 3. issue-rate table incl. pairs, latency and copy contention
timeout 120 python scripts/umma_microbench.py > gpurun_out/${TAG}_umma_microbench.json 2> gpurun_out/${TAG}_umma_microbench.txt; tail -12 gpurun_out/${TAG}_umma_microbench.txt
# 4. Mission Bay (BASELINE.json configs[3], mip renderer, fp32 CUDA path) has no throughput number yet
timeout 300 python - <<'PY' > gpurun_out/${TAG}_mission_bay.json 2> gpurun_out/${TAG}_mission_bay.err
import json, time, torch
from switch_nerf_b200 import synthetic as O
from switch_nerf_b200.configs import make_hparams
from switch_nerf_b200.nerf_moe import get_nerf_moe_inner
from switch_nerf_b200.rendering_mip import render_rays
N, S = 13312, 257
hp = make_hparams(num_experts=8, model_chunk_size=212992, coarse_samples=S, fine_samples=S, width=512,
                  nerfmoe_class_name="MipNeRFMoE")
hp.perturb = 0
model = get_nerf_moe_inner(hp, 2048, 3)
model.load_state_dict(O.synthetic_state_dict(num_experts=8, appearance_count=2048, seed=0, gate_scale=4.0, width=512))
model = model.cuda().eval()
rays, idx = O.synthetic_rays(N, 2048, seed=7)
rays[:, 6], rays[:, 7] = 0.01, 10.0
radii = torch.rand(N, 1) * 1.5e-3 + 5e-4
rays, idx, radii = rays.cuda(), idx.cuda(), radii.cuda()
for _ in range(2):
    render_rays(model, rays, radii, idx, hp, True, True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3):
    render_rays(model, rays, radii, idx, hp, True, True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
print(json.dumps({"workload": "mission bay 13312 rays x (256+256) intervals, width 512, fp32 CUDA path", "ms_per_step": dt * 1e3,
                  "samples_per_s": N * 512 / dt}))
PY
cat gpurun_out/${TAG}_mission_bay.json; tail -3 gpurun_out/${TAG}_mission_bay.err
