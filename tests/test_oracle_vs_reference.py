"""CPU, only where /root/reference exists (this container): pin the oracle restatement against the LIVE
unmodified reference on fresh seeds (the committed goldens pin it on fixed seeds everywhere else)."""
import os
import warnings

import pytest
import torch

HAVE_REF = os.path.isdir("/root/reference/switch_nerf")
pytestmark = pytest.mark.skipif(not HAVE_REF, reason="needs /root/reference (not present on the GPU box)")


@pytest.mark.parametrize("E,cf,bpr,seed", [(4, 1.0, True, 21), (8, 0.5, False, 22), (8, 2.0, True, 23)])
def test_render_rays_bit_exact(E, cf, bpr, seed):
    from oracle import ref_shims as R, switch_nerf_oracle as O
    warnings.filterwarnings("ignore")
    R.install_shims()
    from switch_nerf import rendering
    sd = O.synthetic_state_dict(num_experts=E, appearance_count=16, seed=seed, gate_scale=3.0)
    hp = R.make_hparams(num_experts=E, capacity_factor=cf, bpr=bpr, model_chunk_size=2048, coarse_samples=24, fine_samples=16)
    m = R.build_reference_model(hp, appearance_count=16).eval()
    m.load_state_dict(sd)
    rays, idx = O.synthetic_rays(96, 16, seed=seed + 1)
    with torch.no_grad(), R.stable_argsort():
        res, _ = rendering.render_rays(m, None, rays, idx, hp, None, None, True, True, False)
    mine = O.render_rays(sd, O.default_cfg(sd, cf, bpr), rays, idx, coarse_samples=24, fine_samples=16, model_chunk_size=2048)
    for k in ("rgb_fine", "depth_fine", "depth_variance_fine", "gate_loss_coarse", "gate_loss_fine"):
        assert torch.equal(res[k], mine[k]), k
    assert torch.equal(res["moe_gates_fine"], mine["moe_gates_fine"])


def test_extract_critical_matches_route_top1_with_ties():
    from oracle import ref_shims as R, switch_nerf_oracle as O
    from oracle.make_golden import make_gates
    R.install_shims()
    from switch_nerf.modules.tutel_moe_ext.tutel_fast_dispatch import extract_critical
    gates = make_gates(6000, 8, 99, 3.0, 0.3, 0.1)
    with R.stable_argsort():
        (_, idx_s, loc_s, gates_s, cap), l_aux = extract_critical(gates, 1, 1.0, True, True)
    i2, l2, g2, c2, a2 = O.route_top1(gates, 1.0, True)
    assert torch.equal(i2, idx_s[0]) and torch.equal(l2, loc_s[0]) and c2 == cap and float(a2) == float(l_aux)


@pytest.mark.parametrize("cs,fs,far,seed", [(24, 16, 1.0, 61), (20, 0, 2.5, 62)])
def test_bg_branch_vs_live_reference(cs, fs, far, seed):
    """oracle.bg_oracle (bg NeRF, sphere geometry, render_rays with bg_nerf) against the live reference on fresh seeds."""
    from oracle import bg_oracle as B, ref_shims as R, switch_nerf_oracle as O
    from oracle.make_golden_bg import bg_hparams, reference_bg
    warnings.filterwarnings("ignore")
    R.install_shims()
    from switch_nerf import rendering
    sd = O.synthetic_state_dict(num_experts=4, appearance_count=16, seed=seed, gate_scale=3.0)
    hp = bg_hparams(R.make_hparams(num_experts=4, capacity_factor=1.0, bpr=True, model_chunk_size=1500, coarse_samples=cs,
                                   fine_samples=fs), layers=6, skip=3, width=64)
    m = R.build_reference_model(hp, appearance_count=16).eval()
    m.load_state_dict(sd)
    bg = reference_bg(hp, 16, seed + 2)
    bg_sd = {k: v.detach() for k, v in bg.state_dict().items()}
    rays, idx = O.synthetic_rays(80, 16, seed=seed + 1)
    rays[:, 7] = far
    c, r = torch.tensor([0.01, 0.02, -0.03]), torch.tensor([1.05, 0.95, 1.1])
    with torch.no_grad(), R.stable_argsort():
        ref, present = rendering.render_rays(m, bg, rays, idx, hp, c, r, True, True, True)
    mine = B.render_rays_with_bg(sd, O.default_cfg(sd, 1.0, True), bg_sd, dict(layers=6, skip_layer=3), rays, idx, c, r,
                                 coarse_samples=cs, fine_samples=fs, model_chunk_size=1500)
    assert bool(present) == bool(int(mine["_present"])) and present
    typ = "fine" if fs > 0 else "coarse"
    for k in (f"rgb_{typ}", f"depth_{typ}", f"fg_rgb_{typ}", f"bg_rgb_{typ}", f"bg_depth_{typ}", f"bg_lambda_{typ}",
              f"depth_variance_{typ}", "gate_loss_coarse"):
        tol = 1e-6 * max(1.0, float(ref[k].abs().max()))
        assert float((ref[k] - mine[k]).abs().max()) <= tol, k
    # model and geometry alone, bit for bit
    x = torch.cat([torch.nn.functional.normalize(torch.randn(500, 3), dim=-1), torch.rand(500, 1),
                   torch.nn.functional.normalize(torch.randn(500, 3), dim=-1), torch.randint(0, 16, (500, 1)).float()], 1)
    with torch.no_grad():
        assert torch.equal(bg(x), B.bg_nerf_forward(x, bg_sd, layers=6, skip_layer=3))
    z = torch.rand(80, 9)
    p_ref, d_ref = rendering._depth2pts_outside(rays[:, None, 0:3], rays[:, None, 3:6], z, c, r, False, False)
    p, d = B.depth2pts_outside(rays[:, None, 0:3], rays[:, None, 3:6], z, c, r)
    assert torch.equal(p_ref, p) and torch.equal(d_ref, d)
    assert torch.equal(rendering._intersect_sphere(rays[:, 0:3], rays[:, 3:6], c, r), B.intersect_sphere(rays[:, 0:3], rays[:, 3:6], c, r))
