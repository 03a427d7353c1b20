"""tcgen05.mma issue-rate table on this GPU (see snb_umma_microbench in include/switch_nerf_b200.h).
   python scripts/umma_microbench.py -> JSON lines; floor = max(M,128)*N/256 clocks per K=16 instruction."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from switch_nerf_b200 import _lib as L  # noqa: E402


def main():
    torch.zeros(1, device="cuda")
    reps = 2048
    rows = []
    for N in (64, 128, 256):
        for ts in (0, 1):
            for flags in (0, 1, 2, 3):
                out = (C.c_uint64 * 6)()
                L.check(L.lib().snb_umma_microbench(N, ts, flags, reps, out, L.stream_handle()))
                clk = out[0] / reps
                rows.append({"N": N, "A": "tmem" if ts else "smem", "bulk_copy": bool(flags & 1), "ldtm": bool(flags & 2),
                             "clk_per_mma": round(clk, 1), "floor": N // 2, "copies_8k": int(out[1]),
                             "copy_B_per_clk": round(out[1] * 8192 / max(out[0], 1), 1),
                             "ldtm_B_per_clk": round(sum(out[2:6]) * 32 * 32 * 4 / max(out[0], 1), 1)})
    for N in (128, 256):            # CTA pairs: M = 256 across a 2-CTA cluster (cta_group::2)
        for ts in (0, 1):
            for reps_p in (1, 16, 2048):
                out = (C.c_uint64 * 6)()
                L.check(L.lib().snb_umma_microbench(N, ts, 4, reps_p, out, L.stream_handle()))
                rows.append({"N": N, "A": "tmem" if ts else "smem", "cta_group": 2, "reps": reps_p,
                             "clk_total": int(out[0]), "clk_per_mma": round(out[0] / reps_p, 1), "floor": N // 2})
    for N, ts in ((256, 0), (256, 1)):          # latency: a single instruction / a short burst, one CTA
        for reps_p in (1, 4, 16):
            out = (C.c_uint64 * 6)()
            L.check(L.lib().snb_umma_microbench(N, ts, 0, reps_p, out, L.stream_handle()))
            rows.append({"N": N, "A": "tmem" if ts else "smem", "cta_group": 1, "reps": reps_p,
                         "clk_total": int(out[0]), "clk_per_mma": round(out[0] / reps_p, 1), "floor": N // 2})
    print(json.dumps(rows))
    for r in rows:
        print(r, file=sys.stderr)


if __name__ == "__main__":
    main()
