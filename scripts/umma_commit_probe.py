"""What slows tcgen05.mma inside the fused kernels relative to the bare issue loop (138 clk, TS form, N = 256)?
Adds the kernels' per-slice habits to the microbenchmark one at a time.  python scripts/umma_commit_probe.py"""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from switch_nerf_b200 import _lib as L

torch.zeros(1, device="cuda")
CASES = [("bare", 0), ("commit/4", 0x200), ("A stride 16", 0x400), ("wait(set)/4", 0x1000), ("commit/4 + wait/4", 0x1200),
         ("B ring 4x32K", 0x2000 | (2 << 4)), ("acc=0 per 16", 0x4000), ("all", 0x7600 | (2 << 4)),
         ("all + copies", 0x7601 | (2 << 4)), ("all + copies + ld readers", 0x7603 | (2 << 4))]
rows = []
for ts in (1, 0):
    for name, fl in CASES:
        if not ts and (fl & 0x400):
            continue
        out = (C.c_uint64 * 6)()
        L.check(L.lib().snb_umma_microbench(256, ts, fl, 20000, out, L.stream_handle()))
        rows.append({"A": "tmem" if ts else "smem", "case": name, "clk_per_mma": round(out[0] / 20000, 1)})
        print(rows[-1], flush=True)
print(json.dumps(rows))
