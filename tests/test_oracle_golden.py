"""CPU: the oracle restatement reproduces the committed golden fixtures, which were written by the
UNMODIFIED reference (oracle/make_golden.py).  Bit-exact in fp32."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import switch_nerf_oracle as O
from oracle.make_golden import ROUTE_CASES, make_gates, model_inputs
from tests.util import sd_checksum
from tests.util import (CUDA_MIP_GOLDENS, CUDA_MODEL_GOLDENS, bf16_contract_check, cuda_golden_case, golden_sd,
                        load_golden)


def _sha(t):
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


@pytest.mark.parametrize("case", ROUTE_CASES, ids=[c[0] for c in ROUTE_CASES])
def test_route_golden(case):
    name, S, E, cf, bpr, seed, temp, tie, sat = case
    g = load_golden("route_cases.npz")
    gates = make_gates(S, E, seed, temp, tie, sat)
    assert _sha(gates) == str(g[f"{name}/sha"][3]), "gate generator drifted"
    idx, loc, gv, cap, l_aux = O.route_top1(gates, cf, bpr)
    assert cap == int(g[f"{name}/cap"][0])
    assert _sha(idx) == str(g[f"{name}/sha"][0])
    assert _sha(loc) == str(g[f"{name}/sha"][1])
    assert _sha(gv) == str(g[f"{name}/sha"][2])
    assert float(l_aux) == float(g[f"{name}/l_aux"][0])
    if S <= 8192:
        assert np.array_equal(idx.numpy(), g[f"{name}/idx"]) and np.array_equal(loc.numpy(), g[f"{name}/loc"])


def test_capacity_formula():
    # tutel_fast_dispatch.py:210-211 ; Python int() truncation of a double product
    assert O.capacity_of(131072, 8, 1.0) == 16384
    assert O.capacity_of(131072, 8, 0.5) == 8192
    assert O.capacity_of(131072, 8, 2.0) == 32768
    assert O.capacity_of(5000, 8, 0.5) == 312
    assert O.capacity_of(7, 8, 1.0) == 1
    assert O.capacity_of(10, 1, 0.7) == int(0.7 * 10)


@pytest.mark.parametrize("tag", ["e4_cf1_bpr_fp32", "e8_cf05_nobpr_fp32", "e8_cf1_bpr_fp32_s777", "e4_nobatch_fp32"])
def test_model_golden_fp32(tag):
    g = load_golden(f"model_{tag}.npz")
    E, cf, bpr, S, seed, gs, count, nobatch, _ = g["params"]
    sd = golden_sd(g)
    x = torch.from_numpy(g["x"])
    assert torch.equal(x, model_inputs(int(S), int(count), int(seed) + 100))
    cfg = O.default_cfg(sd, float(cf), bool(bpr), moe_no_batch=bool(nobatch))
    out, ex = O.nerf_moe_forward(x, sd, cfg, mode="fp32")
    assert np.array_equal(out.numpy(), g["outputs"])
    assert np.array_equal(ex["idx"].numpy(), g["idx"])
    assert np.array_equal(ex["loc"].numpy(), g["loc"])
    assert float(ex["l_aux"]) == float(g["l_aux"][0])


def test_model_golden_bf16():
    """The bf16 rounding map of the oracle vs the reference under torch.autocast('cpu', bf16):
    outputs are bf16 (ulp 2^-8 near 1), so agreement is 'within one output ulp, almost always exact'."""
    g = load_golden("model_e8_cf1_bpr_bf16cpu.npz")
    sd = golden_sd(g)
    cfg = O.default_cfg(sd, 1.0, True)
    out, ex = O.nerf_moe_forward(torch.from_numpy(g["x"]), sd, cfg, mode="bf16", flavor="cpu")
    d = np.abs(out.numpy() - g["outputs"])
    same_route = ex["idx"].numpy() == g["idx"]
    assert same_route.mean() > 0.995
    assert d[same_route].max() <= 2 ** -7 + 1e-6      # <= 2 bf16 ulps of a value in [0.5, 1)
    assert (d > 1e-3).mean() < 0.03
    assert d.mean() < 1e-4


@pytest.mark.parametrize("tag", CUDA_MODEL_GOLDENS + CUDA_MIP_GOLDENS)
def test_oracle_cuda_flavor_pinned_to_reference_on_b200(tag):
    """The rounding map the tcgen05 path is held to (mode="bf16", flavor="cuda") vs the UNMODIFIED reference run on a
    B200 under torch.autocast("cuda", bfloat16) (oracle/make_golden_cuda.py; incl. one full 131072-row Building chunk
    with the benchmark's weights): the bf16 contract of tests/util.py with >= 99.9 % identical routing (the host CPU's GEMM accumulation order moves a few near-ties)."""
    c = cuda_golden_case(tag)
    cfg = O.default_cfg(c["sd"], c["cf"], c["bpr"], moe_no_batch=c["no_batch"], mip=c["mip"])
    out, ex = O.nerf_moe_forward(c["x"], c["sd"], cfg, mode="bf16", flavor="cuda")
    kept = None if c["no_batch"] else ex["loc"] < c["cap"]
    if not c["no_batch"]:
        assert ex["capacity"] == c["cap"]
    bf16_contract_check(out, ex["idx"], kept, c, route_min=0.999)
    assert abs(float(ex["l_aux"]) - float(c["g"]["l_aux"][0])) < 2e-3 * abs(float(c["g"]["l_aux"][0]))


@pytest.mark.parametrize("tag", ["config1", "config1_coarse_only", "ragged_chunks"])
def test_render_golden(tag):
    g = load_golden(f"render_{tag}.npz")
    E, cf, bpr, n_rays, cs, fs, chunk, seed, gs, count = g["params"]
    sd = golden_sd(g)
    cfg = O.default_cfg(sd, float(cf), bool(bpr))
    res = O.render_rays(sd, cfg, torch.from_numpy(g["rays"]), torch.from_numpy(g["image_indices"]),
                        coarse_samples=int(cs), fine_samples=int(fs), model_chunk_size=int(chunk))
    typ = "fine" if fs > 0 else "coarse"
    for k in (f"rgb_{typ}", f"depth_{typ}", f"depth_variance_{typ}", "gate_loss_coarse"):
        assert np.array_equal(res[k].numpy(), g[k]), k
    assert np.array_equal(res["moe_gates_coarse"].numpy().astype(np.int32), g["moe_gates_coarse"])


@pytest.mark.parametrize("tag", ["mip_w256", "mip_mission_bay_w512"])
def test_render_mip_golden(tag):
    """oracle restatement of rendering_mip.py + MipNeRFMoE vs the reference-written fixture (bit-exact)."""
    g = load_golden(f"render_{tag}.npz")
    E, width, n_rays, cs, fs, chunk, seed, gs, count = g["params"]
    sd = O.synthetic_state_dict(num_experts=int(E), appearance_count=int(count), seed=int(seed), gate_scale=float(gs), width=int(width))
    res = O.render_rays_mip(sd, O.default_cfg(sd, 1.0, True, mip=True), torch.from_numpy(g["rays"]), torch.from_numpy(g["radii"]),
                            torch.from_numpy(g["image_indices"]), coarse_samples=int(cs), fine_samples=int(fs), model_chunk_size=int(chunk))
    for k in ("rgb_coarse", "rgb_fine", "depth_fine", "depth_variance_fine", "gate_loss_coarse", "gate_loss_fine"):
        assert np.array_equal(res[k].numpy(), g[k]), k


def test_oracle_gradients_vs_reference_golden():
    """Next scope row (backward): the oracle's autograd through the restated path reproduces the parameter
    gradients the UNMODIFIED reference produced for one training-style step (oracle/make_golden_grad.py): identical
    loss and bit-identical gradients (fp32, CPU).  This also pins which paths are cut: the fine samples come from DETACHED coarse weights
    (rendering.py:240)."""
    from oracle import make_golden_grad as G
    g = load_golden("grad_config1.npz")
    loss, grads = G.oracle_grads()
    assert loss == float(g["loss"][0])
    scale = float(g["scale"][0])
    names = [k[len("sample/"):] for k in g if k.startswith("sample/")]
    assert len(names) == 32 and set(names) == set(grads)
    for k in names:
        mine = G.sample_of(grads[k])
        ref = torch.from_numpy(g["sample/" + k])
        assert torch.equal(mine, ref), k              # same torch ops in the same order -> bit-identical on CPU
        st = g["stats/" + k]
        assert float(grads[k].double().abs().sum()) == st[1] and float(grads[k].abs().max()) == st[2], k


def test_backward_plan_model_chunk_equals_autograd():
    """oracle/backward_plan.py (the stage-by-stage backward the CUDA kernels will implement) == autograd of the pinned
    forward restatement, for one model chunk with drops (cf 0.5) and BPR: every parameter gradient."""
    from oracle import backward_plan as B
    from switch_nerf_b200 import synthetic as SY
    torch.manual_seed(0)
    sd = SY.synthetic_state_dict(num_experts=4, appearance_count=8, seed=21, gate_scale=3.0)
    S = 700
    x = torch.cat([torch.rand(S, 3) - 0.5, torch.nn.functional.normalize(torch.randn(S, 3), dim=1),
                   torch.randint(0, 8, (S, 1)).float()], 1)
    for cf, bpr in ((0.5, True), (1.0, False), (2.0, True)):
        cfg = O.default_cfg(sd, cf, bpr)
        sdg = {k: v.clone().double().requires_grad_(True) for k, v in sd.items()}
        R = torch.randn(S, 4).double()
        c_aux = 0.37
        orig_float = torch.Tensor.float
        torch.Tensor.float = lambda self, *a, **k: self.double()      # run the restatement in fp64 (tight comparison)
        try:
            out, ex = O.nerf_moe_forward(x.double(), sdg, cfg, "fp32")
            ((out * R).sum() + c_aux * ex["l_aux"]).backward()
            sd64 = {k: v.double() for k, v in sd.items()}
            torch.set_default_dtype(torch.float64)
            try:
                grads = B.model_chunk_backward(x.double(), sd64, cfg, R, c_aux)
            finally:
                torch.set_default_dtype(torch.float32)
        finally:
            torch.Tensor.float = orig_float
        assert (ex["loc"] >= ex["capacity"]).any() or cf >= 2.0          # the cf 0.5 / 1.0 cases really drop samples
        assert set(grads) == {k for k, v in sdg.items() if v.grad is not None}
        for k, g in grads.items():
            ref = sdg[k].grad
            err = float((g - ref).abs().max())
            assert err <= 1e-9 * max(1.0, float(ref.abs().max())), (cf, bpr, k, err)


def test_backward_plan_composite_and_merge_equal_autograd():
    from oracle import backward_plan as B
    g = torch.Generator().manual_seed(3)
    N, Sf, Sc = 37, 9, 13
    zf = torch.sort(torch.rand(N, Sf, generator=g, dtype=torch.float64) * 0.9 + 0.05, -1)[0]
    zc = torch.sort(torch.rand(N, Sc, generator=g, dtype=torch.float64) * 0.9 + 0.05, -1)[0]
    raw_f = torch.rand(N, Sf, 4, generator=g, dtype=torch.float64).requires_grad_(True)
    raw_c = torch.rand(N, Sc, 4, generator=g, dtype=torch.float64).requires_grad_(True)
    last = 1e10 * torch.ones(N, 1, dtype=torch.float64)
    z_all, order = torch.sort(torch.cat([zf, zc], -1), -1)
    rgbs = torch.gather(torch.cat([raw_f[..., :3], raw_c[..., :3]], 1), 1, order.unsqueeze(-1).expand(-1, -1, 3))
    sig = torch.gather(torch.cat([raw_f[..., 3] * 30, raw_c[..., 3] * 30], 1), 1, order)
    comp = O.composite(z_all, rgbs, sig, last)
    d_rgb = torch.randn(N, 3, generator=g, dtype=torch.float64)
    (comp["rgb"] * d_rgb).sum().backward()
    d_rgbs, d_sig = B.composite_backward(z_all, rgbs.detach(), sig.detach(), last, d_rgb)
    (drf, dsf), (drc, dsc) = B.merge_backward(order, Sf, d_rgbs, d_sig)
    for mine, ref in ((drf, raw_f.grad[..., :3]), (dsf * 30, raw_f.grad[..., 3]), (drc, raw_c.grad[..., :3]), (dsc * 30, raw_c.grad[..., 3])):
        assert float((mine - ref).abs().max()) <= 1e-10 * max(1.0, float(ref.abs().max()))


# ----------------------------------------------------------------------------- row f3: background branch
def _bg_sd(layers, skip, width, softplus, seed, count):
    """Weights of the reference's bg model for a fixture: the mirror's constructor draws the same values under the same
    seed (checked against the stored checksum)."""
    from torch import nn
    from switch_nerf_b200.nerf import NeRF, ShiftedSoftplus
    torch.manual_seed(seed)
    bg = NeRF(12, 4, layers, [skip], width, 48, False, count, 3, 4, ShiftedSoftplus() if softplus else nn.ReLU())
    with torch.no_grad():
        bg.sigma.bias += 1.5                       # oracle.make_golden_bg.BG_SIGMA_BIAS
    return {k: v.detach() for k, v in bg.state_dict().items()}


@pytest.mark.parametrize("tag", ["l8_w256_softplus", "l4_w64_relu"])
def test_bg_oracle_model_pinned_to_reference(tag):
    from oracle import bg_oracle as B
    g = load_golden(f"bg_model_{tag}.npz")
    S, layers, skip, width, softplus, seed, count = (int(v) for v in g["params"])
    sd = _bg_sd(layers, skip, width, softplus, seed, count)
    ck = float(g["sd_checksum"][0])
    assert abs(sd_checksum(sd) - ck) < 1e-6 * ck
    x = torch.from_numpy(g["x"])
    out = B.bg_nerf_forward(x, sd, layers=layers, skip_layer=skip, shifted_softplus=bool(softplus))
    out_n = B.bg_nerf_forward(x, sd, layers=layers, skip_layer=skip, shifted_softplus=bool(softplus),
                              sigma_noise=torch.from_numpy(g["noise"]))
    assert float((out - torch.from_numpy(g["out"])).abs().max()) <= 1e-6
    assert float((out_n - torch.from_numpy(g["out_noise"])).abs().max()) <= 1e-6


def test_bg_oracle_sphere_geometry_pinned_to_reference():
    from oracle import bg_oracle as B
    g = load_golden("bg_sphere_s24.npz")
    rays, z = torch.from_numpy(g["rays"]), torch.from_numpy(g["z"])
    for name, c, r in (("scaled", torch.from_numpy(g["center"]), torch.from_numpy(g["radius"])), ("unit", None, None)):
        far = B.intersect_sphere(rays[:, 0:3], rays[:, 3:6], c, r)
        pts, real = B.depth2pts_outside(rays[:, None, 0:3], rays[:, None, 3:6], z, c, r)
        assert torch.equal(far, torch.from_numpy(g[f"fg_far_{name}"]))
        assert torch.equal(pts, torch.from_numpy(g[f"pts_{name}"]))
        assert torch.equal(real, torch.from_numpy(g[f"depth_real_{name}"]))
    bad = rays[:1].clone()
    bad[0, :3] = torch.tensor([5.0, 0.0, 0.0])
    with pytest.raises(Exception, match="bounded by the unit sphere"):
        B.intersect_sphere(bad[:, 0:3], bad[:, 3:6], None, None)


@pytest.mark.parametrize("tag", ["fine", "coarse_only", "none_leave"])
def test_bg_oracle_render_pinned_to_reference(tag):
    """oracle.bg_oracle.render_rays_with_bg == the unmodified reference's render_rays(nerf, bg_nerf, ...) on every key."""
    from oracle import bg_oracle as B
    g = load_golden(f"bg_render_{tag}.npz")
    E, n_rays, cs, fs, chunk, seed, gs, count, far = g["params"]
    sd = O.synthetic_state_dict(num_experts=int(E), appearance_count=int(count), seed=int(seed), gate_scale=float(gs))
    assert abs(sd_checksum(sd) - float(g["sd_checksum"][0])) < 1e-6 * float(g["sd_checksum"][0])
    bg_sd = _bg_sd(8, 4, 256, True, int(seed) + 2, int(count))
    assert abs(sd_checksum(bg_sd) - float(g["bg_checksum"][0])) < 1e-6 * float(g["bg_checksum"][0])
    res = B.render_rays_with_bg(sd, O.default_cfg(sd, 1.0, True), bg_sd, dict(layers=8, skip_layer=4),
                                torch.from_numpy(g["rays"]), torch.from_numpy(g["image_indices"]),
                                torch.from_numpy(g["center"]), torch.from_numpy(g["radius"]),
                                coarse_samples=int(cs), fine_samples=int(fs), model_chunk_size=int(chunk))
    assert int(res["_present"]) == int(g["present"][0])
    typ = "fine" if fs > 0 else "coarse"
    keys = [k for k in g if k.endswith(f"_{typ}") or k.startswith("gate_loss")]
    assert len(keys) >= 9
    for k in keys:
        a, b = res[k], torch.from_numpy(g[k])
        # same torch CPU ops in the same order: equal up to the last bit of a differently associated sum
        tol = 1e-6 * max(1.0, float(b.abs().max()))
        assert a.shape == b.shape and float((a - b).abs().max()) <= tol, (k, float((a - b).abs().max()))
