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
