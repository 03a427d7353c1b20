"""One-shot hardware probe for the TS-form UMMA (A operand in tensor memory): with B = I the selftest returns
D = A as the tensor core interprets the tcgen05.st-written TMEM image, which pins (or reveals) the layout.
   python scripts/ts_probe.py  -> JSON"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from switch_nerf_b200 import _lib as L  # noqa: E402


def main():
    out = {}
    for K in (16, 32, 64, 128, 256):
        N = K
        a = (torch.arange(128 * K, dtype=torch.float32).reshape(128, K) % 251 + 1).bfloat16().cuda()   # exact in bf16
        b = torch.eye(N, K).bfloat16().cuda()
        d = torch.zeros(128, N, dtype=torch.float32, device="cuda")
        rc = L.lib().snb_umma_selftest(L.ptr(a), L.ptr(b), N, K, L.ptr(d), 4, L.stream_handle())
        torch.cuda.synchronize()
        ok = bool(torch.equal(d, a.float()))
        entry = {"rc": rc, "identity_roundtrip": ok}
        if not ok:
            # where did A[r, k] end up?  report the column permutation seen on row 0 and the row permutation on col 0
            af, df = a.float().cpu(), d.cpu()
            entry["row0_expected"] = af[0, :16].tolist()
            entry["row0_got"] = df[0, :16].tolist()
            entry["col0_expected"] = af[:8, 0].tolist()
            entry["col0_got"] = df[:8, 0].tolist()
            entry["rows_equal"] = int((af == df).all(1).sum())
        out[f"K{K}"] = entry
    print(json.dumps(out))


if __name__ == "__main__":
    main()
