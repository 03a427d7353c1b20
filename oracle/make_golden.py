"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz by running the UNMODIFIED
reference (/root/reference/switch_nerf, imported through oracle/ref_shims.py) on
seeded synthetic inputs.  Run here (the GPU box has no /root/reference):

    python -m oracle.make_golden

Weights are NOT stored: they are regenerated from the recorded seed with
`oracle.switch_nerf_oracle.synthetic_state_dict` (torch CPU generators are
deterministic for a fixed torch build) and loaded into the reference model with
`load_state_dict`; a checksum of the weights is stored so a silent generator
change is caught by the tests.
"""
import hashlib
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shims as R            # noqa: E402
from oracle import switch_nerf_oracle as O   # noqa: E402

warnings.filterwarnings("ignore")
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sd_checksum(sd):
    return float(sum(float(v.double().abs().sum()) for v in sd.values()))


def digest(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def make_gates(S, E, seed, temperature=1.0, tie_frac=0.0, saturate_frac=0.0):
    """Softmax gates with optional exact ties of the row max across rows (BPR tie
    order) and saturated rows (max gate == 1.0f)."""
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(S, E, generator=g) * temperature
    if saturate_frac > 0:
        n = int(S * saturate_frac)
        rows = torch.randperm(S, generator=g)[:n]
        logits[rows, torch.randint(0, E, (n,), generator=g)] += 200.0
    if tie_frac > 0:
        n = int(S * tie_frac)
        rows = torch.randperm(S, generator=g)[:n]
        if n > 0:
            logits[rows] = logits[rows[0]].clone()      # identical rows -> identical max gate
            perm = torch.stack([torch.randperm(E, generator=g) for _ in range(n)])
            logits[rows] = torch.gather(logits[rows], 1, perm)   # ... routed to different experts
    return torch.softmax(logits, dim=1)


ROUTE_CASES = [
    # name, S, E, cf, bpr, seed, temperature, tie_frac, saturate_frac
    ("s4096_e4_cf1_bpr", 4096, 4, 1.0, True, 11, 1.0, 0.0, 0.0),
    ("s4096_e4_cf1_nobpr", 4096, 4, 1.0, False, 12, 1.0, 0.0, 0.0),
    ("s5000_e8_cf05_bpr_ties", 5000, 8, 0.5, True, 13, 2.0, 0.2, 0.05),
    ("s8192_e8_cf2_bpr_sat", 8192, 8, 2.0, True, 14, 4.0, 0.0, 0.3),
    ("s1_e8_cf1_bpr", 1, 8, 1.0, True, 15, 1.0, 0.0, 0.0),
    ("s7_e8_cf1_bpr", 7, 8, 1.0, True, 16, 1.0, 0.0, 0.0),
    ("s3001_e16_cf1_nobpr", 3001, 16, 1.0, False, 17, 3.0, 0.1, 0.0),
    ("s131072_e8_cf1_bpr", 131072, 8, 1.0, True, 18, 3.0, 0.05, 0.02),
    ("s131072_e8_cf05_nobpr", 131072, 8, 0.5, False, 19, 1.5, 0.0, 0.0),
]


def golden_route():
    R.install_shims()
    from switch_nerf.modules.tutel_moe_ext.tutel_fast_dispatch import extract_critical
    out = {}
    for name, S, E, cf, bpr, seed, temp, tie, sat in ROUTE_CASES:
        gates = make_gates(S, E, seed, temp, tie, sat)
        with R.stable_argsort():
            (nE, idx_s, loc_s, gates_s, cap), l_aux = extract_critical(gates, 1, cf, True, bpr)
        idx, loc, gv = idx_s[0], loc_s[0], gates_s[0]
        # the contract requires argmax tie-break; make sure the case has no exact row-max ties
        top2 = torch.topk(gates, min(2, E), dim=1).values
        assert E == 1 or bool((top2[:, 0] > top2[:, 1]).all()), name
        # pin the restatement too
        i2, l2, g2, c2, a2 = O.route_top1(gates, cf, bpr)
        assert torch.equal(i2, idx) and torch.equal(l2, loc) and c2 == cap and torch.equal(g2, gv), name
        assert float((a2 - l_aux).abs()) == 0.0
        out[f"{name}/params"] = np.array([S, E, cf, int(bpr), seed, temp, tie, sat], dtype=np.float64)
        out[f"{name}/cap"] = np.array([cap], dtype=np.int64)
        out[f"{name}/l_aux"] = l_aux.numpy().reshape(1)
        if S <= 8192:
            out[f"{name}/idx"], out[f"{name}/loc"], out[f"{name}/gate"] = idx.numpy(), loc.numpy(), gv.numpy()
        out[f"{name}/sha"] = np.array([digest(idx), digest(loc), digest(gv), digest(gates)])
        print(name, "cap", cap, "dropped", int((loc >= cap).sum()), "l_aux", float(l_aux))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "route_cases.npz"), **out)


def model_inputs(S, appearance_count, seed):
    g = torch.Generator().manual_seed(seed)
    xyz = (torch.rand(S, 3, generator=g) - 0.5) * 1.6
    d = torch.nn.functional.normalize(torch.randn(S, 3, generator=g), dim=-1)
    a = torch.randint(0, appearance_count, (S, 1), generator=g).float()
    return torch.cat([xyz, d, a], 1)


def golden_model(tag, E, cf, bpr, S, autocast_bf16=False, no_batch=False, gate_scale=4.0, seed=3):
    appearance_count = 16
    sd = O.synthetic_state_dict(num_experts=E, appearance_count=appearance_count, seed=seed, gate_scale=gate_scale)
    hp = R.make_hparams(num_experts=E, capacity_factor=cf, bpr=bpr, model_chunk_size=4096,
                        coarse_samples=32, fine_samples=32, amp_bf16=autocast_bf16,
                        moe_expert_type="seqexperts" if no_batch else "expertmlp")
    m = R.build_reference_model(hp, appearance_count=appearance_count).eval()
    if no_batch:
        from switch_nerf.models.model_utils import convert_to_seqexperts
        sd_ref = {k.replace("module.", ""): v for k, v in convert_to_seqexperts({k: v.clone() for k, v in sd.items()}).items()}
        m.load_state_dict(sd_ref)
        m.set_no_batch(True)
    else:
        m.load_state_dict(sd)
    x = model_inputs(S, appearance_count, seed + 100)
    with torch.no_grad(), R.stable_argsort():
        if autocast_bf16:
            with torch.autocast("cpu", dtype=torch.bfloat16):
                r = m(x)
        else:
            r = m(x)
    out = r["outputs"].float()
    gates_idx = r["extras"]["moe_gates"][0].view(-1).to(torch.int32)
    l_aux = r["extras"]["moe_loss"].float()
    # restatement must agree (bit-exact in fp32)
    cfg = O.default_cfg(sd, cf, bpr, moe_no_batch=no_batch)
    o2, ex = O.nerf_moe_forward(x, sd, cfg, mode="bf16" if autocast_bf16 else "fp32", flavor="cpu")
    err = float((o2 - out).abs().max())
    print(tag, "oracle-vs-reference max abs", err, "dropped", int((ex["loc"] >= ex["capacity"]).sum()),
          "counts", torch.bincount(ex["idx"].long(), minlength=E).tolist())
    if not autocast_bf16:
        assert err == 0.0
        assert torch.equal(ex["idx"], gates_idx)
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, f"model_{tag}.npz"),
        params=np.array([E, cf, int(bpr), S, seed, gate_scale, appearance_count, int(no_batch), int(autocast_bf16)], dtype=np.float64),
        sd_checksum=np.array([sd_checksum(sd)]), x=x.numpy(), outputs=out.numpy(), idx=gates_idx.numpy(),
        loc=ex["loc"].numpy(), l_aux=l_aux.numpy(), capacity=np.array([ex["capacity"]]),
        gates=ex["gates"].numpy().astype(np.float32))


def golden_render(tag, E, cf, bpr, n_rays, cs, fs, chunk, gate_scale=4.0, seed=5):
    from switch_nerf import rendering
    appearance_count = 16
    sd = O.synthetic_state_dict(num_experts=E, appearance_count=appearance_count, seed=seed, gate_scale=gate_scale)
    hp = R.make_hparams(num_experts=E, capacity_factor=cf, bpr=bpr, model_chunk_size=chunk,
                        coarse_samples=cs, fine_samples=fs)
    m = R.build_reference_model(hp, appearance_count=appearance_count).eval()
    m.load_state_dict(sd)
    rays, idx = O.synthetic_rays(n_rays, appearance_count, seed=seed + 1)
    with torch.no_grad(), R.stable_argsort():
        res, _ = rendering.render_rays(m, None, rays, idx, hp, None, None, True, True, False)
    cfg = O.default_cfg(sd, cf, bpr)
    mine = O.render_rays(sd, cfg, rays, idx, coarse_samples=cs, fine_samples=fs, model_chunk_size=chunk)
    typ = "fine" if fs > 0 else "coarse"
    for k in (f"rgb_{typ}", f"depth_{typ}", f"depth_variance_{typ}", "gate_loss_coarse"):
        assert float((res[k] - mine[k]).abs().max()) == 0.0, k
    save = {k: v.numpy() for k, v in res.items()}
    save["moe_gates_coarse"] = save["moe_gates_coarse"].astype(np.int32)
    if fs > 0:
        save["moe_gates_fine"] = save["moe_gates_fine"].astype(np.int32)
        save["z_fine"] = mine["_z_fine"].numpy()
        save["raw_fine"] = mine["_raw_fine"].numpy()
    save["raw_coarse"] = mine["_raw_coarse"].numpy()
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, f"render_{tag}.npz"),
        params=np.array([E, cf, int(bpr), n_rays, cs, fs, chunk, seed, gate_scale, appearance_count], dtype=np.float64),
        sd_checksum=np.array([sd_checksum(sd)]), rays=rays.numpy(), image_indices=idx.numpy(), **save)
    print("render", tag, {k: tuple(v.shape) for k, v in res.items()})


def golden_render_mip(tag, E, width, n_rays, cs, fs, chunk, gate_scale=3.0, seed=9):
    """rendering_mip.render_rays + MipNeRFMoE (Mission Bay topology when width=512), eval mode."""
    R.install_shims()
    from switch_nerf import rendering_mip
    appearance_count = 16
    sd = O.synthetic_state_dict(num_experts=E, appearance_count=appearance_count, seed=seed, gate_scale=gate_scale, width=width)
    hp = R.make_hparams(num_experts=E, model_chunk_size=chunk, coarse_samples=cs, fine_samples=fs, width=width,
                        nerfmoe_class_name="MipNeRFMoE")
    hp.perturb = 0
    m = R.build_reference_model(hp, appearance_count=appearance_count).eval()
    m.load_state_dict(sd)
    rays, idx = O.synthetic_rays(n_rays, appearance_count, seed=seed + 1)
    g = torch.Generator().manual_seed(seed + 2)
    radii = torch.rand(n_rays, 1, generator=g) * 1.5e-3 + 5e-4          # radii ~ U(5e-4, 2e-3)
    with torch.no_grad(), R.stable_argsort():
        res, _ = rendering_mip.render_rays(m, rays, radii, idx, hp, True, True)
    mine = O.render_rays_mip(sd, O.default_cfg(sd, 1.0, True, mip=True), rays, radii, idx, coarse_samples=cs,
                             fine_samples=fs, model_chunk_size=chunk)
    for k in ("rgb_coarse", "rgb_fine", "depth_fine", "depth_variance_fine", "gate_loss_coarse", "gate_loss_fine"):
        assert float((res[k] - mine[k]).abs().max()) == 0.0, k
    save = {k: v.numpy() for k, v in res.items()}
    save["moe_gates_coarse"] = save["moe_gates_coarse"].astype(np.int32)
    save["moe_gates_fine"] = save["moe_gates_fine"].astype(np.int32)
    save["z_fine"] = mine["_z_fine"].numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"render_{tag}.npz"),
                        params=np.array([E, width, n_rays, cs, fs, chunk, seed, gate_scale, appearance_count], dtype=np.float64),
                        sd_checksum=np.array([sd_checksum(sd)]), rays=rays.numpy(), radii=radii.numpy(),
                        image_indices=idx.numpy(), **save)
    print("render", tag, {k: tuple(v.shape) for k, v in res.items()})


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(8)
    golden_route()
    golden_model("e4_cf1_bpr_fp32", 4, 1.0, True, 4096)
    golden_model("e8_cf05_nobpr_fp32", 8, 0.5, False, 5000)
    golden_model("e8_cf1_bpr_fp32_s777", 8, 1.0, True, 777)
    golden_model("e4_nobatch_fp32", 4, 1.0, False, 3000, no_batch=True)
    golden_model("e8_cf1_bpr_bf16cpu", 8, 1.0, True, 4096, autocast_bf16=True)
    # BASELINE.json configs[0]: 256 rays x 64 samples, 4 experts, cf=1.0
    golden_render("config1", 4, 1.0, True, 256, 32, 32, 4096)
    golden_render("config1_coarse_only", 4, 1.0, True, 256, 64, 0, 4096)
    golden_render("ragged_chunks", 8, 1.0, True, 100, 17, 9, 1000)
    golden_render_mip("mip_w256", 4, 256, 128, 33, 33, 3000)
    golden_render_mip("mip_mission_bay_w512", 8, 512, 64, 33, 33, 1500)


if __name__ == "__main__":
    main()
