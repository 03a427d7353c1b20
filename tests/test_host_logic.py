"""CPU: host-side mirror of the reference interface (state_dict layout, init parity, error behaviour)."""
import os

import pytest
import torch

from switch_nerf_b200 import synthetic as S
from switch_nerf_b200.configs import make_hparams
from switch_nerf_b200.nerf_moe import NeRFMoE, get_nerf_moe_inner

HAVE_REF = os.path.isdir("/root/reference/switch_nerf")


def test_state_dict_keys_and_shapes_match_reference_layout():
    hp = make_hparams(num_experts=8)
    m = get_nerf_moe_inner(hp, 2048, 3)
    sd = m.state_dict()
    ref = S.synthetic_state_dict(num_experts=8, appearance_count=2048)
    assert set(sd) == set(ref)
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    assert sd["layers.0.experts.0.weights.3"].shape == (8, 256, 256)      # [E, in, out]  
    assert sd["layers.0.experts.0.bias.0"].shape == (8, 1, 256)
    assert sd["layers.0.gates.0.wg.weight"].shape == (8, 256)
    assert sd["layers.2.fcs.0.weight"].shape == (128, 331)


@pytest.mark.skipif(not HAVE_REF, reason="needs /root/reference")
def test_init_is_bit_identical_to_reference_under_same_seed():
    from oracle import ref_shims as R
    hp = make_hparams(num_experts=4)
    ref = R.build_reference_model(hp, appearance_count=16, seed=7)
    torch.manual_seed(7)
    mine = get_nerf_moe_inner(make_hparams(num_experts=4), 16, 3)
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys()) or set(a) == set(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_forward_shape_error_and_no_cpu_path():
    m = get_nerf_moe_inner(make_hparams(num_experts=4), 16, 3)
    with pytest.raises(Exception, match="Unexpected input shape"):
        m(torch.zeros(3, 5))
    from switch_nerf_b200._lib import SnbError
    with pytest.raises(SnbError):
        m(torch.zeros(3, 7))             # parameters / inputs on CPU: refuse, never fall back


def test_unsupported_topology_is_rejected():
    hp = make_hparams(num_experts=4)
    hp.use_moe_external_gate = False
    with pytest.raises(NotImplementedError):
        get_nerf_moe_inner(hp, 16, 3)


def test_set_no_batch_toggles_moe_layers():
    m = get_nerf_moe_inner(make_hparams(num_experts=4), 16, 3)
    assert m.layers["0"].moe_no_batch is False
    m.set_no_batch(True)
    assert m.layers["0"].moe_no_batch is True and m.route_opts().no_batch == 1
