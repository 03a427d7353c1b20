"""Per-SM weight ingest vs tcgen05.mma rate: the issue-rate microbenchmark (TS form, N = 256) with four bulk copies of
8 / 16 / 32 KB kept in flight -- the fused kernels' weight ring -- reporting the landed bytes per clock and the MMA rate
under that traffic.  python scripts/umma_ingest.py"""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from switch_nerf_b200 import _lib as L

torch.zeros(1, device="cuda")
rows = []
for ts, chip in ((1, 0), (0, 0), (1, 0x100)):
    for sz in (0, 1, 2):
        for reps in (20000,):
            out = (C.c_uint64 * 6)()
            L.check(L.lib().snb_umma_microbench(256, ts, 1 | (sz << 4) | chip, reps, out, L.stream_handle()))
            rows.append({"A": "tmem" if ts else "smem", "ctas": 148 if chip else 1, "copy_kb": 8 << sz, "in_flight": 4, "clk_per_mma": round(out[0] / reps, 1),
                         "copies": int(out[1]), "ingest_B_per_clk": round(out[1] * (8192 << sz) / max(out[0], 1), 1)})
            print(rows[-1])
print(json.dumps(rows))
