"""Stand-alone timing of snb_route_select (k_pack_top1 + k_select) on one Building-size chunk."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from switch_nerf_b200 import _lib as L

S, E = 131072, 8
g = torch.Generator().manual_seed(1)
spread = float(os.environ.get("SPREAD", "1.0"))
gates = torch.softmax(torch.randn(S, E, generator=g) * spread, 1).cuda()
lib = L.lib()
idx = torch.zeros(S, dtype=torch.int32, device="cuda"); loc = torch.zeros_like(idx)
gv = torch.zeros(S, device="cuda"); counts = torch.zeros(E, dtype=torch.int32, device="cuda")
cap = torch.zeros(1, dtype=torch.int32, device="cuda"); la = torch.zeros(1, device="cuda")
nb = lib.snb_route_select_workspace_bytes(S)
ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
def run(loc_t):
    L.check(lib.snb_route_select(L.ptr(gates), S, E, 1.0, 1, 0, None, L.ptr(loc_t) if loc_t is not None else None, None,
                                 L.ptr(counts), L.ptr(cap), L.ptr(la), L.ptr(ws), nb, L.stream_handle()))
for taps in (None, loc):
    for _ in range(3): run(taps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run(taps)
    e1.record(); torch.cuda.synchronize()
    print("taps" if taps is not None else "no taps", "pack+select us:", e0.elapsed_time(e1) / 20 * 1e3, "counts", counts.tolist())
