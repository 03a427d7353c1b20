"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/bg_*.npz by running the UNMODIFIED reference
(/root/reference/switch_nerf through oracle/ref_shims.py) on seeded synthetic inputs: the background NeRF
(models/nerf.py:75-191 with xyz_dim = 4), the sphere geometry (rendering.py:497-570) and render_rays with
bg_nerf (rendering.py:15-196).  Run here (the GPU box has no /root/reference):

    python -m oracle.make_golden_bg

Weights are not stored: the foreground state_dict comes from `synthetic_state_dict(seed)`, the background model is
the reference constructor under `torch.manual_seed(seed)` (+ a fixed sigma-bias shift so the background is not
transparent); the mirror in switch_nerf_b200/nerf.py constructs the same modules in the same order, and a checksum
of the weights is stored so a drift of either is caught.
"""
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shims as R            # noqa: E402
from oracle import switch_nerf_oracle as O   # noqa: E402
from oracle.make_golden import GOLDEN_DIR, sd_checksum  # noqa: E402

warnings.filterwarnings("ignore")

BG_SIGMA_BIAS = 1.5


def bg_hparams(hp, layers=8, skip=4, width=256):
    hp.layers, hp.skip_layers, hp.bg_layer_dim, hp.bg_use_cfg, hp.ckpt_path = layers, [skip], width, False, None
    hp.expertmlp2seqexperts = False
    return hp


def reference_bg(hp, count, seed):
    R.install_shims()
    from switch_nerf.models.model_utils import get_bg_nerf
    torch.manual_seed(seed)
    bg = get_bg_nerf(hp, count).eval()
    with torch.no_grad():
        bg.sigma.bias += BG_SIGMA_BIAS
    return bg


def golden_bg_model(tag, S, layers, skip, width, softplus, seed=21, count=16):
    hp = bg_hparams(R.make_hparams(), layers, skip, width)
    hp.shifted_softplus = softplus
    bg = reference_bg(hp, count, seed)
    g = torch.Generator().manual_seed(seed + 1)
    p = torch.nn.functional.normalize(torch.randn(S, 3, generator=g), dim=-1)
    inv = torch.rand(S, 1, generator=g)
    d = torch.nn.functional.normalize(torch.randn(S, 3, generator=g), dim=-1)
    a = torch.randint(0, count, (S, 1), generator=g).float()
    x = torch.cat([p, inv, d, a], 1)
    noise = torch.randn(S, 1, generator=g) * 0.5
    with torch.no_grad():
        out = bg(x)
        out_noise = bg(x, sigma_noise=noise)
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"bg_model_{tag}.npz"),
                        params=np.array([S, layers, skip, width, int(softplus), seed, count], dtype=np.float64),
                        sd_checksum=np.array([sd_checksum(bg.state_dict())]), x=x.numpy(), noise=noise.numpy(),
                        out=out.numpy(), out_noise=out_noise.numpy())
    print("bg model", tag, tuple(out.shape), float(out[:, 3].mean()))


def sphere_case(n_rays, seed):
    rays, idx = O.synthetic_rays(n_rays, 16, seed=seed)
    center = torch.tensor([0.02, -0.01, 0.03])
    radius = torch.tensor([1.1, 0.9, 1.0])
    return rays, idx, center, radius


def golden_sphere(tag, n_rays=300, S=24, seed=31):
    R.install_shims()
    from switch_nerf import rendering
    rays, _, center, radius = sphere_case(n_rays, seed)
    g = torch.Generator().manual_seed(seed + 1)
    z = torch.sort(torch.rand(n_rays, S, generator=g), -1)[0]
    z[:, 0], z[:, -1] = 0.0, 1.0
    o, d = rays[:, 0:3], rays[:, 3:6]
    save = {}
    for name, (c, r) in {"scaled": (center, radius), "unit": (None, None)}.items():
        far = rendering._intersect_sphere(o, d, c, r)
        pts, real = rendering._depth2pts_outside(o.view(-1, 1, 3), d.view(-1, 1, 3), z, c, r, False, False)
        save[f"fg_far_{name}"], save[f"pts_{name}"], save[f"depth_real_{name}"] = far.numpy(), pts.numpy(), real.numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"bg_sphere_{tag}.npz"), rays=rays.numpy(), z=z.numpy(),
                        center=center.numpy(), radius=radius.numpy(), **save)
    print("sphere", tag, {k: v.shape for k, v in save.items()})


def golden_render_bg(tag, E, n_rays, cs, fs, chunk, seed=41, gate_scale=4.0, count=16, far=1.0):
    from switch_nerf import rendering
    sd = O.synthetic_state_dict(num_experts=E, appearance_count=count, seed=seed, gate_scale=gate_scale)
    hp = bg_hparams(R.make_hparams(num_experts=E, capacity_factor=1.0, bpr=True, model_chunk_size=chunk, coarse_samples=cs,
                                   fine_samples=fs))
    m = R.build_reference_model(hp, appearance_count=count).eval()
    m.load_state_dict(sd)
    bg = reference_bg(hp, count, seed + 2)
    rays, idx, center, radius = sphere_case(n_rays, seed + 1)
    rays[:, 7] = far
    with torch.no_grad(), R.stable_argsort():
        res, present = rendering.render_rays(m, bg, rays, idx, hp, center, radius, True, True, True)
    o, d = rays[:, 0:3], rays[:, 3:6]
    fg_far = torch.maximum(rendering._intersect_sphere(o, d, center, radius), rays[:, 6])
    save = {k: v.numpy() for k, v in res.items() if not k.startswith("moe_gates")}
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"bg_render_{tag}.npz"),
                        params=np.array([E, n_rays, cs, fs, chunk, seed, gate_scale, count, far], dtype=np.float64),
                        sd_checksum=np.array([sd_checksum(sd)]), bg_checksum=np.array([sd_checksum(bg.state_dict())]),
                        rays=rays.numpy(), image_indices=idx.numpy(), center=center.numpy(), radius=radius.numpy(),
                        present=np.array([int(present)]), n_with_bg=np.array([int((rays[:, 7] > fg_far).sum())]), **save)
    print("render bg", tag, present, int((rays[:, 7] > fg_far).sum()), {k: tuple(v.shape) for k, v in save.items()})


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(8)
    golden_bg_model("l8_w256_softplus", 3000, 8, 4, 256, True)
    golden_bg_model("l4_w64_relu", 1500, 4, 2, 64, False)
    golden_sphere("s24")
    golden_render_bg("fine", 4, 200, 32, 32, 4096)              # some rays end inside the sphere (far = 1.0), some leave it
    golden_render_bg("coarse_only", 4, 128, 48, 0, 1000, far=3.0)  # every ray continues into the background; ragged chunks
    golden_render_bg("none_leave", 4, 64, 16, 16, 4096, far=0.5)   # no ray reaches the sphere: bg branch empty


if __name__ == "__main__":
    main()
