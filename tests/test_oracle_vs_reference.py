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
