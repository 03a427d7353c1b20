"""GPU box: the tcgen05 path vs the fresh CUDA-autocast goldens of the unmodified reference (oracle/make_golden_cuda.py).
    python scripts/compare_cuda_golden.py gpurun_out/golden_cuda > gpurun_out/compare_cuda_golden.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import switch_nerf_oracle as O                      # noqa: E402
from oracle.make_golden_cuda import bench_chunk_x, bench_inputs  # noqa: E402
from switch_nerf_b200 import synthetic as SY                     # noqa: E402
from tests.util import bf16_contract_stats, make_model           # noqa: E402

d = sys.argv[1]
out = {}
for f in sorted(os.listdir(d)):
    if not (f.startswith("model_") and f.endswith(".npz")):
        continue
    g = dict(np.load(os.path.join(d, f)))
    p = g["params"]
    E, cf, bpr, S, seed, gs, count, nobatch, width, mip = int(p[0]), float(p[1]), bool(p[2]), int(p[3]), int(p[4]), float(p[5]), int(p[6]), bool(p[7]), int(p[9]), bool(p[10])
    if width != 256 or mip:
        continue
    if "bench_chunk" in f:
        sd, rays, idx = bench_inputs()
        x = bench_chunk_x(rays, idx, 257, S)
    else:
        sd = SY.synthetic_state_dict(num_experts=E, appearance_count=count, seed=seed, gate_scale=gs)
        x = torch.from_numpy(g["x"])
    model, _ = make_model(sd, cf, bpr, no_batch=nobatch, precision="bf16")
    with torch.no_grad():
        r = model(x.cuda(), return_debug=True)
    torch.cuda.synchronize()
    o = r["outputs"].cpu()
    idx_m = r["extras"]["moe_gates"][0].view(-1).cpu()
    loc_m = r["extras"]["debug_loc"].cpu()
    cap = int(g["capacity"][0])
    st = bf16_contract_stats(o, torch.from_numpy(g["outputs"]), idx_m, torch.from_numpy(g["idx"]),
                             None if nobatch else loc_m < cap, None if nobatch else torch.from_numpy(g["loc"]) < cap)
    # for scale: the oracle's CUDA-flavoured bf16 map on the CPU against the same golden
    if S <= 8192:
        o2, ex = O.nerf_moe_forward(x, sd, O.default_cfg(sd, cf, bpr, moe_no_batch=nobatch), mode="bf16", flavor="cuda")
        st["oracle_cuda_flavor"] = bf16_contract_stats(o2, torch.from_numpy(g["outputs"]), ex["idx"], torch.from_numpy(g["idx"]))
    out[f] = st
    print(f, json.dumps(st), file=sys.stderr)
print(json.dumps(out, indent=1))
