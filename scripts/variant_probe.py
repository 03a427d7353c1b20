"""Measurement aid: time one Building model chunk (131072 rows, bf16 path) with the library given by SNB_LIB
(default: the regular build).  Prints ms per chunk and the per-phase split; variants may compute garbage."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from switch_nerf_b200 import _lib as L
if os.environ.get("SNB_LIB"):
    L.LIB_PATH = os.path.abspath(os.environ["SNB_LIB"])
from switch_nerf_b200 import synthetic as O
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
with torch.no_grad():
    for _ in range(5):
        model(x)
    torch.cuda.synchronize()
    reps = 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        model(x)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{os.environ.get('SNB_LIB', 'default')}: {ms:.3f} ms per 131072-row chunk = {S / ms / 1e3:.1f} M samples/s (serial chunk: front + route + back)")
