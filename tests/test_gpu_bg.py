"""Row f3: background NeRF + sphere parametrisation against the UNMODIFIED reference's outputs
(tests/golden/bg_*.npz, written by oracle/make_golden_bg.py from models/nerf.py:75-191 and rendering.py:15-196, 497-570
on CPU in fp32).  Tolerances: fp32 kernels vs fp32 torch on another device -- 1e-5 for the geometry (angles go through
asin/sin/cos), 2e-5 for the 8-layer MLP outputs, TOL = 1e-3 (BASELINE.json north_star) for composited rays."""
import ctypes as C

import numpy as np
import pytest
import torch
from torch import nn

from oracle import ref_shims as R
from oracle.make_golden_bg import BG_SIGMA_BIAS, bg_hparams
from tests.util import golden_sd, load_golden, make_model, sd_checksum

pytestmark = pytest.mark.gpu
TOL = 1e-3


def make_bg(layers, skip, width, softplus, seed, count):
    """The mirror under the seed the reference model was built with (same module creation order -> same weights)."""
    from switch_nerf_b200.nerf import NeRF, ShiftedSoftplus
    torch.manual_seed(seed)
    bg = NeRF(12, 4, layers, [skip], width, 48, False, count, 3, 4, ShiftedSoftplus() if softplus else nn.ReLU()).eval()
    with torch.no_grad():
        bg.sigma.bias += BG_SIGMA_BIAS
    return bg


@pytest.mark.parametrize("tag", ["l8_w256_softplus", "l4_w64_relu"])
def test_bg_model_vs_reference_golden(built_lib, tag):
    g = load_golden(f"bg_model_{tag}.npz")
    S, layers, skip, width, softplus, seed, count = (int(v) for v in g["params"])
    bg = make_bg(layers, skip, width, softplus, seed, count)
    ck = float(g["sd_checksum"][0])
    assert abs(sd_checksum(bg.state_dict()) - ck) < 1e-6 * ck, "constructor RNG order differs from models/nerf.py"
    bg = bg.cuda()
    x = torch.from_numpy(g["x"]).cuda()
    with torch.no_grad():
        out = bg(x).cpu().numpy()
        out_noise = bg(x, sigma_noise=torch.from_numpy(g["noise"]).cuda()).cpu().numpy()
        empty = bg(x[:0])
    assert empty.shape == (0, 4)
    assert np.abs(out - g["out"]).max() <= 2e-5, np.abs(out - g["out"]).max()
    assert np.abs(out_noise - g["out_noise"]).max() <= 2e-5
    assert np.abs(out_noise[:, :3] - out[:, :3]).max() == 0.0          # the noise only enters sigma
    # parameter update is seen by the next call (version-tracked re-upload)
    with torch.no_grad():
        bg.rgb.bias += 0.25
        out2 = bg(x).cpu().numpy()
    expect = 1 / (1 + np.exp(-(np.log(g["out"][:, :3] / (1 - g["out"][:, :3])) + 0.25)))
    assert np.abs(out2[:, :3] - expect).max() <= 1e-4
    with pytest.raises(Exception):
        bg(x[:, :5])


def test_bg_model_rejects_what_is_not_built(built_lib):
    from switch_nerf_b200 import _lib as L
    from switch_nerf_b200.nerf import NeRF, ShiftedSoftplus
    for kw in (dict(rgb_dim=12), dict(affine=True), dict(dirs=0), dict(xyz_dim=3)):
        with pytest.raises(L.SnbError):
            NeRF(12, kw.get("dirs", 4), 8, [4], 64, 48, kw.get("affine", False), 8, kw.get("rgb_dim", 3), kw.get("xyz_dim", 4),
                 ShiftedSoftplus())
    with pytest.raises(L.SnbError):
        NeRF(12, 4, 8, [4], 64, 48, False, 8, 3, 4, ShiftedSoftplus())(torch.zeros(4, 8))       # CPU tensor: no CPU path


@pytest.mark.parametrize("name", ["scaled", "unit"])
def test_sphere_geometry_vs_reference_golden(built_lib, name):
    """snb_intersect_sphere / snb_depth2pts_outside vs rendering._intersect_sphere / _depth2pts_outside."""
    from switch_nerf_b200 import _lib as L
    lib = L.lib()
    g = load_golden("bg_sphere_s24.npz")
    rays, z = torch.from_numpy(g["rays"]).cuda(), torch.from_numpy(g["z"]).cuda()
    c = torch.from_numpy(g["center"]).cuda() if name == "scaled" else None
    r = torch.from_numpy(g["radius"]).cuda() if name == "scaled" else None
    N, S = z.shape
    far = torch.empty(N, device="cuda")
    bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    pts, real = torch.empty(N, S, 4, device="cuda"), torch.empty(N, S, device="cuda")
    st = L.stream_handle()
    L.check(lib.snb_intersect_sphere(L.ptr(rays), N, L.ptr(c), L.ptr(r), L.ptr(far), L.ptr(bad), st))
    L.check(lib.snb_depth2pts_outside(L.ptr(rays), L.ptr(c), L.ptr(r), L.ptr(z), N, S, L.ptr(pts), L.ptr(real), st))
    torch.cuda.synchronize()
    assert int(bad.item()) == 0
    assert np.abs(far.cpu().numpy() - g[f"fg_far_{name}"]).max() <= 1e-5
    assert np.abs(pts.cpu().numpy() - g[f"pts_{name}"]).max() <= 1e-5
    ref_real = g[f"depth_real_{name}"]
    # depth_real = cos(theta) / (z + 1e-8) + d1 reaches 1e8 at z = 0: relative there
    err = np.abs(real.cpu().numpy() - ref_real) / np.maximum(1.0, np.abs(ref_real))
    assert err.max() <= 1e-5, err.max()
    assert torch.equal(pts[..., 3], z)
    # a camera outside the sphere is reported (the reference raises)
    out_ray = rays[:1].clone()
    out_ray[0, :3] = torch.tensor([5.0, 0.0, 0.0])
    out_ray[0, 3:6] = torch.tensor([0.0, 1.0, 0.0])
    L.check(lib.snb_intersect_sphere(L.ptr(out_ray), 1, L.ptr(c), L.ptr(r), L.ptr(far), L.ptr(bad), st))
    assert int(bad.item()) == 1
    assert lib.snb_intersect_sphere(L.ptr(rays), N, L.ptr(c), None, L.ptr(far), None, st) == (0 if c is None else 1)    # SNB_EINVAL: centre without radius


@pytest.mark.parametrize("tag", ["fine", "coarse_only", "none_leave"])
def test_render_with_bg_vs_reference_golden(built_lib, tag):
    """render_rays(nerf, bg_nerf, ..., sphere_center, sphere_radius, get_bg_fg_rgb=True) vs the reference: every result
    key, the rays-present flag, fp32 foreground."""
    from switch_nerf_b200.rendering import render_rays
    g = load_golden(f"bg_render_{tag}.npz")
    E, n_rays, cs, fs, chunk, seed, gs, count, far = g["params"]
    gg = dict(g)
    gg["params"] = np.array([E, 1.0, 1, n_rays, cs, fs, chunk, seed, gs, count])
    model, hp = make_model(golden_sd(gg), 1.0, True, False, "fp32")
    bg_hparams(hp)
    hp.coarse_samples, hp.fine_samples, hp.model_chunk_size = int(cs), int(fs), int(chunk)
    bg = make_bg(8, 4, 256, True, int(seed) + 2, int(count))
    ck = float(g["bg_checksum"][0])
    assert abs(sd_checksum(bg.state_dict()) - ck) < 1e-6 * ck
    bg = bg.cuda()
    rays, idx = torch.from_numpy(g["rays"]).cuda(), torch.from_numpy(g["image_indices"]).cuda()
    c, r = torch.from_numpy(g["center"]).cuda(), torch.from_numpy(g["radius"]).cuda()
    rays0 = rays.clone()
    with torch.no_grad():
        res, present = render_rays(model, bg, rays, idx, hp, c, r, True, True, True, debug_taps=True)
    torch.cuda.synchronize()
    assert torch.equal(rays, rays0), "render_rays must not modify the caller's rays"
    assert bool(present) == bool(g["present"][0])
    if present:
        assert res["_rays_with_bg"].numel() == int(g["n_with_bg"][0])
    typ = "fine" if fs > 0 else "coarse"
    keys = [k for k in g if k.endswith(f"_{typ}") and not k.startswith("gate_loss")]
    assert set(keys) <= set(res), set(keys) - set(res)
    assert len(keys) >= 8
    for k in keys:
        a, b = res[k].cpu().numpy(), g[k]
        assert a.shape == b.shape, k
        if "depth" in k:
            # bg depths are 1/z-like (depth_real up to 1e8 at z = 0, weighted by tiny weights): relative to the magnitude
            err = (np.abs(a - b) / np.maximum(1.0, np.abs(b))).max()
        else:
            err = np.abs(a - b).max()
        assert err <= TOL, (k, err)
    assert np.abs(res["gate_loss_coarse"].cpu().numpy() - g["gate_loss_coarse"]).max() <= 1e-4
    # the bg contribution is really there: where rays continue, rgb differs from the foreground-only composite
    if present:
        d = (res[f"rgb_{typ}"] - res[f"fg_rgb_{typ}"]).abs().sum(-1)
        assert float(d[res["_rays_with_bg"]].min()) >= 0 and float(d.max()) > 1e-3
        mask = torch.ones(rays.shape[0], dtype=torch.bool, device="cuda")
        mask[res["_rays_with_bg"]] = False
        assert float(d[mask].max() if mask.any() else 0.0) == 0.0


def test_render_with_bg_bf16_foreground_and_training_noise(built_lib):
    """bf16 foreground (the tcgen05 path) + fp32 background stays within one bf16 output ulp of the fp32 render; training
    mode (perturb, sigma noise) gives finite, in-range colours; gradients through the bg branch are refused loudly."""
    from switch_nerf_b200.rendering import render_rays
    g = load_golden("bg_render_fine.npz")
    E, n_rays, cs, fs, chunk, seed, gs, count, far = g["params"]
    gg = dict(g)
    gg["params"] = np.array([E, 1.0, 1, n_rays, cs, fs, chunk, seed, gs, count])
    sd = golden_sd(gg)
    bg = make_bg(8, 4, 256, True, int(seed) + 2, int(count)).cuda()
    rays, idx = torch.from_numpy(g["rays"]).cuda(), torch.from_numpy(g["image_indices"]).cuda()
    c, r = torch.from_numpy(g["center"]).cuda(), torch.from_numpy(g["radius"]).cuda()
    model, hp = make_model(sd, 1.0, True, False, "bf16")
    bg_hparams(hp)
    hp.coarse_samples, hp.fine_samples, hp.model_chunk_size = int(cs), int(fs), int(chunk)
    with torch.no_grad():
        res, present = render_rays(model, bg, rays, idx, hp, c, r, True, True, False)
    assert present
    err = np.abs(res["rgb_fine"].cpu().numpy() - g["rgb_fine"])
    assert err.max() <= 2.0 ** -8 and err.mean() <= 5e-4, (err.max(), err.mean())
    model.train(), bg.train()
    hp.use_sigma_noise, hp.sigma_noise_std, hp.perturb = True, 0.5, 1.0
    import warnings
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res, _ = render_rays(model, bg, rays, idx, hp, c, r, True, True, False, seed=5)
    assert all(torch.isfinite(v.float()).all() for v in res.values())
    assert float(res["rgb_fine"].min()) >= 0 and float(res["rgb_fine"].max()) <= 1 + 1e-4
    with pytest.raises(NotImplementedError):
        render_rays(model, bg, rays, idx, hp, c, r, True, True, False)
