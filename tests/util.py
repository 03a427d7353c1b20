"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

from oracle import ref_shims as R
from oracle import switch_nerf_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sd_checksum(sd):
    return float(sum(float(v.double().abs().sum()) for v in sd.values()))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def golden_sd(g, width=256):
    """Re-generate the weights a fixture was produced with and check them against its checksum."""
    p = g["params"]
    if len(p) == 9:      # model_* fixture
        E, seed, gate_scale, count = int(p[0]), int(p[4]), float(p[5]), int(p[6])
    else:                # render_* fixture
        E, seed, gate_scale, count = int(p[0]), int(p[7]), float(p[8]), int(p[9])
    sd = O.synthetic_state_dict(num_experts=E, appearance_count=count, seed=seed, gate_scale=gate_scale, width=width)
    assert abs(sd_checksum(sd) - float(g["sd_checksum"][0])) < 1e-6 * float(g["sd_checksum"][0]), \
        "weight generator drifted from the one the golden fixture was made with"
    return sd


def make_model(sd, capacity_factor=1.0, bpr=True, no_batch=False, precision="fp32", device="cuda",
               moe_return_gates=True):
    """switch_nerf_b200 NeRFMoE holding `sd` (reference state_dict layout)."""
    from switch_nerf_b200.nerf_moe import get_nerf_moe_inner
    E = sd["layers.0.gates.0.wg.weight"].shape[0]
    width = sd["layers.0.gates.0.wg.weight"].shape[1]
    count = sd["embedding_a.weight"].shape[0]
    hp = R.make_hparams(num_experts=E, capacity_factor=capacity_factor, bpr=bpr, width=width,
                        amp_bf16=(precision == "bf16"), moe_return_gates=moe_return_gates)
    model = get_nerf_moe_inner(hp, count, 3)
    missing = model.load_state_dict(sd)
    assert not missing.missing_keys and not missing.unexpected_keys
    model = model.to(device).eval()
    model.set_no_batch(no_batch)
    return model, hp


def bf16_ulp(x):
    """Spacing of bf16 numbers at |x| (8 significant bits)."""
    e = torch.floor(torch.log2(x.abs().clamp_min(2.0 ** -126)))
    return torch.pow(2.0, e - 7)


def bf16_contract_stats(out, ref_out, idx, ref_idx, kept=None, ref_kept=None):
    """Per-sample [rgb, sigma] of a bf16 path vs the UNMODIFIED reference's CUDA-autocast output (tests/golden/*bf16cuda*):
    routing agreement, and on the samples both sides routed identically the distribution of |d rgb| in bf16 output ulps
    (sigmoid(rgb) is a bf16 value on both sides) and the sigma error (fp32 softplus of a bf16 pre-activation)."""
    same = idx.view(-1).long() == ref_idx.view(-1).long()
    if kept is not None and ref_kept is not None:
        same &= kept.view(-1) == ref_kept.view(-1)
    d_rgb = (out[same, :3] - ref_out[same, :3]).abs()
    ulps = (d_rgb / bf16_ulp(ref_out[same, :3])).round()
    d_sig = (out[same, 3] - ref_out[same, 3]).abs()
    n = max(int(ulps.numel()), 1)
    return {"route_agree": float(same.float().mean()),
            "rgb_ulp0": float((ulps == 0).sum()) / n, "rgb_ulp_le1": float((ulps <= 1).sum()) / n,
            "rgb_ulp_le2": float((ulps <= 2).sum()) / n, "rgb_max_abs": float(d_rgb.max()) if n else 0.0,
            "rgb_frac_le_1e-3": float((d_rgb <= 1e-3).float().mean()),
            "sigma_max_abs": float(d_sig.max()), "sigma_frac_le_1e-3": float((d_sig <= 1e-3).float().mean()),
            "sigma_max_rel": float((d_sig / ref_out[same, 3].abs().clamp_min(1e-2)).max()),
            "mean_abs_all": float((out - ref_out).abs().mean())}


# ----------------------------------------------------------------------------------------------------------------
# bf16 contract of the tcgen05 path, stated against the UNMODIFIED reference run on a B200 under
# torch.autocast("cuda", bfloat16) (tests/golden/model_*_bf16cuda.npz, written by oracle/make_golden_cuda.py):
#   C1  routing: expert id and kept/dropped status equal the reference's on >= 99.8 % of the samples (the fp32 gate sees
#       bf16 GEMM outputs whose fp32 accumulation order differs; only near-ties can flip);
#   C2  rgb, on samples routed identically: EVERY value within one bf16 output ulp of the reference (sigmoid output is a
#       bf16 number on both sides; one ulp <= 2^-8), >= 98 % bit-identical -- i.e. <= 1e-3 abs wherever the
#       reference's own rounding did not flip;
#   C3  sigma, same samples: |d sigma| <= 1e-3 + 2^-7 * sigma_ref for every sample (two bf16 ulps of the pre-activation
#       through the fp32 softplus), <= 1e-3 abs on >= 95 %.
# The oracle's flavor="cuda" rounding map is pinned to the same files with the same bounds and >= 99.9 % routing.
# ----------------------------------------------------------------------------------------------------------------
CUDA_MODEL_GOLDENS = ("e8_cf1_bpr", "e8_cf05_nobpr", "e8_cf2_bpr", "e4_cf1_bpr", "e4_nobatch", "bench_chunk")
CUDA_MIP_GOLDENS = ("mip_e4_w256", "mip_e8_w512")


def cuda_golden_case(tag):
    """Inputs of a model_<tag>_bf16cuda.npz fixture, regenerated from its seeds and checked against its checksum."""
    from oracle.make_golden_cuda import bench_chunk_x, bench_inputs
    from switch_nerf_b200 import synthetic as SY
    g = load_golden(f"model_{tag}_bf16cuda.npz")
    p = g["params"]
    c = dict(E=int(p[0]), cf=float(p[1]), bpr=bool(p[2]), S=int(p[3]), seed=int(p[4]), gate_scale=float(p[5]),
             count=int(p[6]), no_batch=bool(p[7]), width=int(p[9]), mip=bool(p[10]), g=g)
    if tag == "bench_chunk":
        sd, rays, idx = bench_inputs()
        x = bench_chunk_x(rays, idx, 257, c["S"])
    else:
        sd = SY.synthetic_state_dict(num_experts=c["E"], appearance_count=c["count"], seed=c["seed"],
                                     gate_scale=c["gate_scale"], width=c["width"])
        x = torch.from_numpy(g["x"])
    assert abs(sd_checksum(sd) - float(g["sd_checksum"][0])) < 1e-6 * float(g["sd_checksum"][0]), \
        "weight generator drifted from the one the golden fixture was made with"
    c.update(sd=sd, x=x, cap=int(g["capacity"][0]), ref_out=torch.from_numpy(g["outputs"]),
             ref_idx=torch.from_numpy(g["idx"]).long(), ref_kept=torch.from_numpy(g["loc"]) < int(g["capacity"][0]))
    return c


def bf16_contract_check(out, idx, kept, case, route_min=0.998):
    """Assert C1-C3 for per-sample outputs [S,4] / expert ids / kept mask against a cuda_golden_case(); returns the stats."""
    ref_out, ref_idx, ref_kept = case["ref_out"], case["ref_idx"], case["ref_kept"]
    st = bf16_contract_stats(out, ref_out, idx, ref_idx, kept, ref_kept)
    same = idx.view(-1).long() == ref_idx.view(-1)
    if kept is not None:
        same &= kept.view(-1) == ref_kept.view(-1)
    d_sig = (out[same, 3] - ref_out[same, 3]).abs()
    st["sigma_worst_over_bound"] = float((d_sig / (1e-3 + 2.0 ** -7 * ref_out[same, 3].abs())).max())
    assert torch.isfinite(out).all()
    assert st["route_agree"] >= route_min, st                                        # C1
    assert st["rgb_ulp_le1"] == 1.0 and st["rgb_ulp0"] >= 0.98, st                  # C2
    assert st["rgb_max_abs"] <= 2.0 ** -8 + 1e-9, st
    assert st["sigma_worst_over_bound"] <= 1.0 and st["sigma_frac_le_1e-3"] >= 0.95, st   # C3
    return st
