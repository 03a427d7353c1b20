"""BASELINE.json configs[4]: capacity-factor sweep 0.5/1.0/2.0 x batch_prioritized_routing on/off at
8192 rays x (257+257) samples, 8 experts -- routing-imbalance throughput curve.  Also sweeps the gate
sharpness (logit temperature via wg scale) to move the max-expert share.  Prints one JSON line per case."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from switch_nerf_b200 import synthetic as O
from switch_nerf_b200.configs import make_hparams
from switch_nerf_b200.nerf_moe import get_nerf_moe_inner
from switch_nerf_b200.rendering import render_rays

N_RAYS, COARSE, FINE, CHUNK, E = 8192, 257, 257, 131072, 8
rays, idx = O.synthetic_rays(N_RAYS, 2048, seed=100)
rays, idx = rays.cuda(), idx.cuda()
rows = []
base_sd = O.synthetic_state_dict(num_experts=E, appearance_count=2048, seed=0, gate_scale=4.0)
r_cpu, _ = O.synthetic_rays(N_RAYS, 2048, seed=100)
pick = torch.randperm(N_RAYS, generator=torch.Generator().manual_seed(7))[:512]
tt = torch.linspace(0, 1, 64)
zz = r_cpu[pick, 6:7] * (1 - tt) + r_cpu[pick, 7:8] * tt
pts = (r_cpu[pick, None, 0:3] + r_cpu[pick, None, 3:6] * zz[..., None]).reshape(-1, 3)
# gate skew: 0 = raw random init (3 experts take ~95 %), 3 / 8 = partially, 60 = fully balanced
for gate_scale in (0, 3, 8, 60):
    sd = O.balance_gate(base_sd, pts, iters=gate_scale) if gate_scale > 0 else base_sd
    for cf in (0.5, 1.0, 2.0):
        for bpr in (True, False):
            hp = make_hparams(num_experts=E, capacity_factor=cf, bpr=bpr, model_chunk_size=CHUNK, coarse_samples=COARSE,
                              fine_samples=FINE, amp_bf16=True, moe_return_gates=True)
            model = get_nerf_moe_inner(hp, 2048, 3)
            model.load_state_dict(sd)
            model = model.cuda().eval()
            for _ in range(3):
                res = render_rays(model, None, rays, idx, hp, None, None, True, True, False)[0]
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                res = render_rays(model, None, rays, idx, hp, None, None, True, True, False)[0]
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            gates = torch.cat([res["moe_gates_coarse"].view(-1), res["moe_gates_fine"].view(-1)])
            share = torch.bincount(gates, minlength=E).float() / gates.numel()
            # dropped fraction from one instrumented chunk
            dropped_n, total_n = 0, 0
            for key in ("moe_gates_coarse", "moe_gates_fine"):      # model chunks restart with every pass
                gk = res[key].view(-1)
                for i in range(0, gk.numel(), CHUNK):
                    c = torch.bincount(gk[i:i + CHUNK], minlength=E)
                    capc = int(cf * ((min(CHUNK, gk.numel() - i) + E - 1) // E))
                    dropped_n += int(torch.clamp(c - capc, min=0).sum()); total_n += int(c.sum())
            dropped = dropped_n / total_n
            row = {"balance_iters": gate_scale, "capacity_factor": cf, "bpr": bpr, "ms_per_step": ms,
                   "msamples_per_s": N_RAYS * (COARSE + FINE) / ms / 1e3, "max_expert_share": float(share.max()),
                   "dropped_fraction": dropped}
            rows.append(row)
            print(json.dumps(row), flush=True)
            model.release()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/cf_sweep.json", "w"), indent=1)
