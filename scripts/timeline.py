"""Debug: dump the per-phase clock marks of CTA 0 of the fused kernels (SNB_TIMELINE=1)."""
import ctypes as C
import os
import sys

os.environ["SNB_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from switch_nerf_b200 import _lib as L, synthetic as O
if os.environ.get("SNB_LIB"):
    L.LIB_PATH = os.path.abspath(os.environ["SNB_LIB"])
from switch_nerf_b200.configs import make_hparams
from switch_nerf_b200.nerf_moe import get_nerf_moe_inner

sd = O.synthetic_state_dict(num_experts=8, appearance_count=2048, seed=0, gate_scale=4.0)
hp = make_hparams(num_experts=8, amp_bf16=True, moe_return_gates=False)
model = get_nerf_moe_inner(hp, 2048, 3)
model.load_state_dict(sd)
model = model.cuda().eval()
S = 131072
g = torch.Generator().manual_seed(5)
x = torch.cat([(torch.rand(S, 3, generator=g) - 0.5) * 1.6, torch.nn.functional.normalize(torch.randn(S, 3, generator=g), dim=-1),
               torch.randint(0, 2048, (S, 1), generator=g).float()], 1).cuda()
for _ in range(3):
    model(x)
torch.cuda.synchronize()
TL_N = 2048
buf = (C.c_uint64 * (5 * TL_N))()
n = L.lib().snb_debug_timeline(buf, 5 * TL_N)
names = ["front/epi", "front/mma", "back/epi", "back/mma"]
for r in range(4):
    marks = [(buf[r * TL_N + i] >> 48, buf[r * TL_N + i] & 0xFFFFFFFFFFFF) for i in range(TL_N) if buf[r * TL_N + i]]
    if not marks:
        continue
    print("==", names[r], len(marks), "marks; showing tiles 2-3")
    # split by tag 1 (tile start)
    starts = [i for i, (t, _) in enumerate(marks) if t == 1]
    for ti in (2, 3):
        if ti + 1 >= len(starts):
            break
        seg = marks[starts[ti]:starts[ti + 1] + 1]
        t0 = seg[0][1]
        print("  tile", ti, "total cycles", seg[-1][1] - t0)
        print("   ", " ".join(f"{t}:{c - t0}" for t, c in seg))

sel = [(buf[4 * TL_N + i] >> 48, buf[4 * TL_N + i] & 0xFFFFFFFFFFFF) for i in range(64) if buf[4 * TL_N + i]]
if sel:
    print("== k_select CTA 0 (1 start, 2 counted, 10+l level chosen, 3 threshold, 4 ordered pass done, 5 end):",
          " ".join(f"{t}:{c - sel[0][1]}" for t, c in sel))
