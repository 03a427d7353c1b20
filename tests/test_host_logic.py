"""CPU: host-side mirror of the reference interface (state_dict layout, init parity, error behaviour)."""
import os

import numpy as np
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


# ---- expert-parallel exchange (csrc/snb_ep.cu): protocol algebra on the CPU ---------------------------------
@pytest.mark.parametrize("world,E,cf,bpr", [(2, 8, 1.0, True), (2, 4, 0.5, True), (4, 8, 2.0, False), (8, 8, 1.0, True)])
def test_expert_parallel_protocol_equals_local(world, E, cf, bpr):
    """routing decided per source rank + row-wise experts => the exchanged result equals the
    all-local one on every rank, every sample is evaluated exactly once by the owner of its expert, dropped
    samples never leave their rank."""
    from oracle import ep_protocol as P
    from oracle import switch_nerf_oracle as O
    g = torch.Generator().manual_seed(world * 100 + E)
    routing, S_max = [], 0
    for r in range(world):
        S = 700 + 37 * r
        gates = torch.softmax(torch.randn(S, E, generator=g) * 2.0, 1)
        idx, loc, gate_val, cap, _ = O.route_top1(gates, cf, bpr)
        payload = torch.randn(S, 3, generator=g).numpy()
        routing.append((idx.numpy().astype(np.int64), loc.numpy().astype(np.int64), cap, payload))
        S_max = max(S_max, S)
    capmax = O.capacity_of(S_max, E, cf)
    row_fn = lambda e, row: (e + 1) * row.sum()
    drop_fn = lambda row: -1.0
    ret = P.exchange(world, E, capmax, routing, row_fn, drop_fn)
    for r, (idx, loc, cap, payload) in enumerate(routing):
        assert sorted(ret[r]) == list(range(len(idx)))
        for s in range(len(idx)):
            want = row_fn(int(idx[s]), payload[s]) if loc[s] < cap else drop_fn(payload[s])
            assert ret[r][s] == want


def test_expert_parallel_owner_map():
    from switch_nerf_b200.expert_parallel import local_experts, owner_of_expert
    assert [owner_of_expert(e, 8, 2) for e in range(8)] == [0, 0, 0, 0, 1, 1, 1, 1]
    assert list(local_experts(3, 8, 8)) == [3] and list(local_experts(1, 8, 2)) == [4, 5, 6, 7]
    with pytest.raises(ValueError):
        owner_of_expert(0, 8, 3)


def test_a2a_init_fails_loudly_without_cuda(built_lib):
    """No CPU fallback for the exchange either: without a CUDA device the group cannot be created."""
    import ctypes as C
    from switch_nerf_b200 import _lib as L
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    h = C.c_void_p()
    rc = L.lib().snb_a2a_init(0, 2, 8, 4096, 1.0, C.byref(h))
    assert rc != 0 and not h.value
    assert L.lib().snb_a2a_init(0, 3, 8, 4096, 1.0, C.byref(h)) != 0      # 8 experts over 3 ranks
    assert b"shard" in L.lib().snb_last_error()


# ---- checkpoint interop (reference models/model_utils.py:12-28, 136-151) -------------------------------------
def test_checkpoint_layouts_roundtrip_and_load():
    from switch_nerf_b200.checkpoint import load_checkpoint, to_expertmlp, to_seqexperts
    sd = S.synthetic_state_dict(num_experts=4, appearance_count=8, seed=2)
    seq = to_seqexperts(sd)
    assert "layers.0.experts.0.experts.3.layers.6.weight" in seq and not any(".weights." in k for k in seq)
    assert seq["layers.0.experts.0.experts.1.layers.2.weight"].shape == (256, 256)
    assert seq["layers.0.experts.0.experts.1.layers.2.bias"].shape == (256,)
    back = to_expertmlp({"module." + k: v for k, v in seq.items()})      # DDP prefix + seqexperts layout
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    hp = make_hparams(num_experts=4)
    a, b, c = (get_nerf_moe_inner(hp, 8, 3) for _ in range(3))
    a.load_state_dict(sd)
    b.load_state_dict({"module." + k: v for k, v in seq.items()})         # pre-hook normalises
    load_checkpoint(c, {"model_state_dict": seq})
    for k, v in a.state_dict().items():
        assert torch.equal(v, b.state_dict()[k]) and torch.equal(v, c.state_dict()[k]), k


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not mounted")
def test_seqexperts_layout_matches_reference_converter():
    from oracle import ref_shims  # noqa: F401  (installs the tutel/timm import shims)
    from switch_nerf.models.model_utils import convert_to_seqexperts
    from switch_nerf_b200.checkpoint import to_seqexperts
    sd = S.synthetic_state_dict(num_experts=4, appearance_count=8, seed=4)
    ref = convert_to_seqexperts({k: v.clone() for k, v in sd.items()})
    ours = to_seqexperts(sd)
    ref = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in ref.items()}
    assert set(ref) == set(ours)
    for k in ref:
        assert torch.equal(ref[k], ours[k]), k


def test_model_copies_do_not_share_the_c_handle_and_moe_layer_finds_its_owner():
    import copy
    import ctypes as C
    from switch_nerf_b200.nerf_moe import _OWNERS
    hp = make_hparams(num_experts=4)
    m = get_nerf_moe_inner(hp, 8, 3)
    m._handle = C.c_void_p(1234)                       # pretend a packed model exists
    try:
        m2 = copy.deepcopy(m)
    finally:
        m._handle = None
    assert m2._handle is None and m2 is not m
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    assert m.layers["0"]._find_owner() is m and m2.layers["0"]._find_owner() is m2
    assert id(m2) in _OWNERS
    with pytest.raises(Exception):                     # no CPU path for the operator either
        m.layers["0"](torch.zeros(4, 256))


def test_bg_nerf_mirror_layout_and_seed_parity():
    """switch_nerf_b200.nerf.NeRF (background model, xyz_dim = 4): the reference's state_dict keys / shapes
    (models/nerf.py:75-150) and, under the same seed, the weights the reference constructor draws -- pinned by the
    checksum oracle/make_golden_bg.py stored from the unmodified reference."""
    from torch import nn
    from switch_nerf_b200.nerf import NeRF, ShiftedSoftplus
    from tests.util import load_golden, sd_checksum
    g = load_golden("bg_model_l8_w256_softplus.npz")
    S_, layers, skip, width, softplus, seed, count = (int(v) for v in g["params"])
    torch.manual_seed(seed)
    bg = NeRF(12, 4, layers, [skip], width, 48, False, count, 3, 4, ShiftedSoftplus())
    with torch.no_grad():
        bg.sigma.bias += 1.5                                            # oracle.make_golden_bg.BG_SIGMA_BIAS
    ck = float(g["sd_checksum"][0])
    assert abs(sd_checksum(bg.state_dict()) - ck) < 1e-6 * ck
    sd = bg.state_dict()
    assert sd["xyz_encodings.0.0.weight"].shape == (256, 100)           # 4 + 4*12*2
    assert sd["xyz_encodings.4.0.weight"].shape == (256, 356)           # skip: [encoded input | hidden]
    assert sd["dir_a_encoding.0.weight"].shape == (128, 256 + 27 + 48)
    assert sd["sigma.weight"].shape == (1, 256) and sd["rgb.weight"].shape == (3, 128)
    assert sd["embedding_a.weight"].shape == (count, 48) and "xyz_encoding_final.bias" in sd
    assert len(sd) == 2 * layers + 9
    if HAVE_REF:
        from oracle import ref_shims as R
        R.install_shims()
        from switch_nerf.models.nerf import NeRF as RefNeRF
        torch.manual_seed(3)
        ref = RefNeRF(12, 4, 8, [4], 64, 48, False, 5, 3, 4, nn.ReLU())
        torch.manual_seed(3)
        mine = NeRF(12, 4, 8, [4], 64, 48, False, 5, 3, 4, nn.ReLU())
        a, b = ref.state_dict(), mine.state_dict()
        assert list(a.keys()) == list(b.keys())
        assert all(torch.equal(a[k], b[k]) for k in a)


def test_render_rays_rejects_foreign_models_before_touching_cuda():
    """The renderer mirror refuses models it cannot hand to the library (no silent torch fallback)."""
    from switch_nerf_b200._lib import SnbError
    from switch_nerf_b200.rendering import render_rays
    hp = make_hparams(num_experts=4)
    m = get_nerf_moe_inner(hp, 16, 3)
    rays = torch.zeros(4, 8)
    with pytest.raises(SnbError):
        render_rays(torch.nn.Linear(3, 3), None, rays, None, hp)
    with pytest.raises(SnbError):
        render_rays(m, torch.nn.Linear(3, 3), rays, None, hp)          # bg_nerf must be the NeRF mirror
    with pytest.raises(SnbError):
        render_rays(m, None, rays, None, hp)                            # CPU rays: no CPU path
    hp.use_cascade = True
    with pytest.raises(NotImplementedError):
        render_rays(m, None, rays, None, hp)


def test_mark_dirty_resets_the_packed_versions():
    from switch_nerf_b200.nerf import NeRF, ShiftedSoftplus
    m = get_nerf_moe_inner(make_hparams(num_experts=4), 16, 3)
    m._packed_versions = ("x",)
    m.mark_dirty()
    assert m._packed_versions is None
    bg = NeRF(12, 4, 4, [2], 64, 48, False, 8, 3, 4, ShiftedSoftplus())
    bg._versions = ("x",)
    bg.mark_dirty()
    assert bg._versions is None
