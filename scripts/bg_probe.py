"""Measurement aid for row f3: render_rays with the background model on a Building-shaped batch (every ray continues into
the background), timed with CUDA events; prints the time of the whole render, of the foreground alone and the
background model's own throughput.  python scripts/bg_probe.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch import nn
from switch_nerf_b200 import synthetic as SY
from switch_nerf_b200.configs import make_hparams
from switch_nerf_b200.nerf import NeRF, ShiftedSoftplus
from switch_nerf_b200.nerf_moe import get_nerf_moe_inner
from switch_nerf_b200.rendering import render_rays

N, CS, FS, COUNT = 2048, 257, 257, 2048
sd = SY.synthetic_state_dict(num_experts=8, appearance_count=COUNT, seed=0, gate_scale=4.0)
hp = make_hparams(num_experts=8, amp_bf16=True, moe_return_gates=False, coarse_samples=CS, fine_samples=FS)
model = get_nerf_moe_inner(hp, COUNT, 3)
model.load_state_dict(sd)
model = model.cuda().eval()
torch.manual_seed(1)
bg = NeRF(12, 4, 8, [4], 256, 48, False, COUNT, 3, 4, ShiftedSoftplus()).cuda().eval()
rays, idx = SY.synthetic_rays(N, COUNT, seed=3)
rays[:, 7] = 3.0                                  # far beyond the unit sphere: every ray has a background segment
rays, idx = rays.cuda(), idx.cuda()


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


with torch.no_grad():
    t_all = timed(lambda: render_rays(model, bg, rays, idx, hp, None, None, True, True, False))
    t_fg = timed(lambda: render_rays(model, None, rays, idx, hp, None, None, True, True, False))
    S = N * (CS // 2 + FS // 2)
    x = torch.cat([torch.nn.functional.normalize(torch.randn(S, 3, device="cuda"), dim=-1), torch.rand(S, 1, device="cuda"),
                   torch.nn.functional.normalize(torch.randn(S, 3, device="cuda"), dim=-1),
                   torch.randint(0, COUNT, (S, 1), device="cuda").float()], 1)
    t_bg = timed(lambda: bg(x))
print(json.dumps({"rays": N, "fg_samples_per_ray": CS + FS, "bg_samples_per_ray": CS // 2 + FS // 2,
                  "render_with_bg_ms": round(t_all, 3), "render_fg_only_ms": round(t_fg, 3),
                  "bg_model_ms": round(t_bg, 3), "bg_model_M_samples_per_s": round(S / t_bg / 1e3, 1),
                  "bg_model_tflops_fp32": round(S * 2 * (100 * 256 + 6 * 256 * 256 + 356 * 256 + 256 * 256 + 256 + 331 * 128 + 128 * 3) / t_bg / 1e9, 1)}))
