"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch-CPU, fp32) of the background branch of the reference's renderer
(SURVEY 8f-3): the background NeRF, the sphere geometry and `render_rays` with `bg_nerf`.  Nothing under
`switch_nerf_b200/` imports this file.

Parity status: PINNED.  `tests/test_oracle_golden.py::test_bg_oracle_*` compare every function with the fixtures
`tests/golden/bg_*.npz` that `oracle/make_golden_bg.py` wrote from the UNMODIFIED reference (bit for bit: the same torch
CPU ops in the same order), and `tests/test_oracle_vs_reference.py` with the reference itself where /root/reference
exists.  All `file:line` citations are relative to /root/reference/switch_nerf/.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from oracle import switch_nerf_oracle as O

Tensor = torch.Tensor


def bg_nerf_forward(x: Tensor, sd: Dict[str, Tensor], *, layers: int, skip_layer: int, pos_xyz_freqs: int = 12,
                    pos_dir_freqs: int = 4, shifted_softplus: bool = True, sigma_noise: Optional[Tensor] = None) -> Tensor:
    """models/nerf.py:137-191 `NeRF.forward` with xyz_dim = 4, rgb_dim = 3, no affine appearance.
    x [S, 8] = [point on the sphere (3), inverse distance (1), direction (3), image index]."""
    input_xyz = O.embedding(x[:, :4], pos_xyz_freqs)                                       # :146 (Embedding, :9-26)
    h = input_xyz
    for i in range(layers):                                                                # :148-151
        if i == skip_layer:
            h = torch.cat([input_xyz, h], -1)
        h = torch.relu(torch.nn.functional.linear(h, sd[f"xyz_encodings.{i}.0.weight"], sd[f"xyz_encodings.{i}.0.bias"]))
    sigma = torch.nn.functional.linear(h, sd["sigma.weight"], sd["sigma.bias"])            # :153
    if sigma_noise is not None:
        sigma = sigma + sigma_noise                                                        # :154-155
    sigma = O.shifted_softplus(sigma) if shifted_softplus else torch.relu(sigma)           # :157
    final = torch.nn.functional.linear(h, sd["xyz_encoding_final.weight"], sd["xyz_encoding_final.bias"])   # :163
    cat = [final, O.embedding(x[:, 4:7], pos_dir_freqs)]                                   # :168
    if "embedding_a.weight" in sd:
        cat.append(sd["embedding_a.weight"][x[:, -1].long()])                              # :171
    d = torch.relu(torch.nn.functional.linear(torch.cat(cat, -1), sd["dir_a_encoding.0.weight"], sd["dir_a_encoding.0.bias"]))
    rgb = torch.nn.functional.linear(d, sd["rgb.weight"], sd["rgb.bias"])                  # :174
    return torch.cat([torch.sigmoid(rgb), sigma], -1)                                      # :191


def _to_unit_sphere(origin: Tensor, direction: Tensor, center: Optional[Tensor], radius: Optional[Tensor]):
    """rendering.py:499-501 / 529-531: the scene sphere becomes the unit sphere."""
    if radius is None:
        return origin, direction
    return (origin - center) / radius, direction / radius


def _closest_approach(origin: Tensor, direction: Tensor) -> Tensor:
    """Ray parameter of the point closest to the sphere centre (rendering.py:508 / 534); negative behind the camera."""
    return -torch.sum(direction * origin, dim=-1) / torch.sum(direction * direction, dim=-1)


def intersect_sphere(rays_o: Tensor, rays_d: Tensor, center: Optional[Tensor], radius: Optional[Tensor]) -> Tensor:
    """rendering.py:497-518: ray parameter at which a ray leaves the unit sphere (after centre / radius)."""
    origin, direction = _to_unit_sphere(rays_o, rays_d, center, radius)
    t_mid = _closest_approach(origin, direction)
    closest = origin + t_mid.unsqueeze(-1) * direction
    inv_len = 1. / torch.norm(direction, dim=-1)
    dist_sq = torch.sum(closest * closest, dim=-1)
    if (dist_sq >= 1.).any():                                    # :513-515, same message
        raise Exception('Not all your cameras are bounded by the unit sphere; please make sure the cameras are normalized properly!')
    half_chord = torch.sqrt(1. - dist_sq) * inv_len
    return t_mid + half_chord


def depth2pts_outside(rays_o: Tensor, rays_d: Tensor, depth: Tensor, center: Optional[Tensor], radius: Optional[Tensor]):
    """rendering.py:521-570 with include_xyz_real = False.  rays_o / rays_d [N,1,3]; depth [N,S] = inverse distance to the
    sphere centre.  Returns the 4-D background points [point on the unit sphere, inverse distance] and the conventional
    depth.  The point is the exit point of the ray rotated towards the ray direction by phi - theta about the axis
    origin x exit (Rodrigues), phi = asin(|closest|), theta = asin(|closest| * depth)."""
    origin, direction = _to_unit_sphere(rays_o, rays_d, center, radius)
    t_mid = _closest_approach(origin, direction)
    closest = origin + t_mid.unsqueeze(-1) * direction
    dist = torch.norm(closest, dim=-1)
    inv_len = 1. / direction.norm(dim=-1)
    half_chord = torch.sqrt(1. - dist * dist) * inv_len                       # note: |closest| squared here, sum of squares in :512
    exit_pt = origin + (t_mid + half_chord).unsqueeze(-1) * direction
    axis = torch.cross(origin, exit_pt, dim=-1)
    axis = axis / (torch.norm(axis, dim=-1, keepdim=True) + 1e-8)
    theta = torch.asin(dist * depth)
    angle = (torch.asin(dist) - theta).unsqueeze(-1)
    cos_a, sin_a = torch.cos(angle), torch.sin(angle)
    along = torch.sum(axis * exit_pt, dim=-1, keepdim=True)
    rotated = exit_pt * cos_a + torch.cross(axis, exit_pt, dim=-1) * sin_a + axis * along * (1. - cos_a)   # :546-548
    rotated = rotated / torch.norm(rotated, dim=-1, keepdim=True)
    depth_real = 1. / (depth + 1e-8) * torch.cos(theta) + t_mid                                            # :552
    return torch.cat((rotated, depth.unsqueeze(-1)), dim=-1), depth_real


def _composite_flip(z_desc: Tensor, rgbs: Tensor, sigmas: Tensor, last_delta: Tensor):
    """rendering.py:436-464 with flip = True: descending depths, deltas z[i] - z[i+1]."""
    deltas = torch.cat([z_desc[..., :-1] - z_desc[..., 1:], last_delta], -1)
    alphas = 1 - torch.exp(-deltas * sigmas)
    T = torch.cumprod(1 - alphas + 1e-8, -1)
    T = torch.cat((torch.ones_like(T[..., 0:1]), T[..., :-1]), dim=-1)
    weights = alphas * T
    return weights, (weights.unsqueeze(-1) * rgbs).sum(dim=1)


def _bg_chunks(pts: Tensor, rays_d: Tensor, idx: Tensor, bg_sd, bg_cfg, chunk: int) -> Tensor:
    n, s = pts.shape[:2]
    x = torch.cat([pts.reshape(-1, 4), rays_d.view(n, 1, 3).expand(n, s, 3).reshape(-1, 3),
                   idx.view(n, 1, 1).expand(n, s, 1).reshape(-1, 1).to(pts.dtype)], 1)
    return torch.cat([bg_nerf_forward(x[i:i + chunk], bg_sd, **bg_cfg) for i in range(0, x.shape[0], chunk)], 0).view(n, s, 4)


def bg_results(rays: Tensor, idx: Tensor, bg_sd, bg_cfg, center, radius, *, coarse_samples: int, fine_samples: int,
               model_chunk_size: int) -> Dict[str, Tensor]:
    """`_get_results(nerf=bg_nerf, flip=True, last_delta=1e10)` (rendering.py:55-77, 199-274, 277-494), eval mode.
    Reference quirks kept: the coarse `depth_real` is not flipped with its samples (:291-294 flip xyz and z only), and the
    resampling pdf pairs the flipped coarse weights with the unflipped bin mid points (:238-241)."""
    n = rays.shape[0]
    rays_o, rays_d = rays[:, None, 0:3], rays[:, None, 3:6]
    sb = coarse_samples // 2
    z = torch.linspace(0, 1, sb).expand(n, sb)
    last = 1e10 * torch.ones(n, 1)
    pts, real_c = depth2pts_outside(rays_o, rays_d, z, center, radius)
    z_c = torch.flip(z, dims=[-1])
    raw_c = _bg_chunks(torch.flip(pts, dims=[-2]), rays[:, 3:6], idx, bg_sd, bg_cfg, model_chunk_size)
    w_c, rgb_c = _composite_flip(z_c, raw_c[..., :3], raw_c[..., 3], last)
    res = {}
    if fine_samples == 0:
        res["rgb_coarse"] = rgb_c
        res["depth_coarse"] = (w_c * real_c).sum(dim=1)
        res["depth_variance_coarse"] = (w_c * (z_c - res["depth_coarse"].unsqueeze(1)).square()).sum(-1)
        return res
    z_mid = 0.5 * (z[:, :-1] + z[:, 1:])
    z_f = O.sample_pdf(z_mid, w_c[:, 1:-1], fine_samples // 2, det=True)
    pts_f, real_f = depth2pts_outside(rays_o, rays_d, z_f, center, radius)
    raw_f = _bg_chunks(pts_f, rays[:, 3:6], idx, bg_sd, bg_cfg, model_chunk_size)
    z_all, order = torch.sort(torch.cat([z_f, z_c], -1), -1, descending=True)                  # :421
    raw_all = torch.gather(torch.cat([raw_f, raw_c], 1), 1, order.unsqueeze(-1).expand(-1, -1, 4))
    real_all = torch.gather(torch.cat([real_f, real_c], 1), 1, order)
    w, rgb = _composite_flip(z_all, raw_all[..., :3], raw_all[..., 3], last)
    res["rgb_fine"] = rgb
    res["depth_fine"] = (w * real_all).sum(dim=1)
    res["depth_variance_fine"] = (w * (z_all - res["depth_fine"].unsqueeze(1)).square()).sum(-1)
    return res


def render_rays_with_bg(sd, cfg, bg_sd, bg_cfg, rays: Tensor, idx: Tensor, center, radius, *, coarse_samples: int,
                        fine_samples: int, model_chunk_size: int) -> Dict[str, Tensor]:
    """rendering.render_rays with bg_nerf (rendering.py:15-196), eval mode, get_bg_fg_rgb = True."""
    n = rays.shape[0]
    near, far = rays[:, 6], rays[:, 7]
    fg_far = torch.maximum(intersect_sphere(rays[:, 0:3], rays[:, 3:6], center, radius), near)     # :34-35
    with_bg = torch.arange(n)[far > fg_far]                                                         # :36
    last_delta = 1e10 * torch.ones(n, 1)
    fg_rays = rays
    bg = {}
    if with_bg.shape[0] > 0:
        last_delta[with_bg, 0] = fg_far[with_bg]                                                    # :42
        fg_rays = rays.clone()
        fg_rays[:, 7] = torch.minimum(far, fg_far)                                                  # :44
        bg = bg_results(rays[with_bg], idx[with_bg], bg_sd, bg_cfg, center, radius, coarse_samples=coarse_samples,
                        fine_samples=fine_samples, model_chunk_size=model_chunk_size)
    # foreground: O.render_rays with last_delta - max(z of the level) where last_delta < 1e10 (:215-216, :250-251)
    fg = _fg_with_last_delta(sd, cfg, fg_rays, idx, last_delta, coarse_samples, fine_samples, model_chunk_size)
    typ = "fine" if fine_samples > 0 else "coarse"
    res = {f"bg_lambda_{typ}": fg["bg_lambda"], "gate_loss_coarse": fg["gate_loss_coarse"]}
    if fine_samples > 0:
        res["gate_loss_fine"] = fg["gate_loss_fine"]
    res[f"depth_variance_{typ}"] = fg["depth_variance"]
    for key in ("rgb", "depth"):                                                                    # :104-146
        val = fg[key]
        bg_val = torch.zeros_like(val)
        if with_bg.shape[0] > 0:
            mult = fg["bg_lambda"][with_bg]
            bg_val[with_bg] = bg[f"{key}_{typ}"] * (mult.unsqueeze(-1) if val.dim() > 1 else mult)
        res[f"fg_{key}_{typ}"], res[f"bg_{key}_{typ}"] = val, bg_val
        res[f"{key}_{typ}"] = val + bg_val if with_bg.shape[0] > 0 else val
    res["_present"] = torch.tensor(int(with_bg.shape[0] > 0))
    return res


def _fg_with_last_delta(sd, cfg, rays, idx, last_delta, cs, fs, chunk):
    """The foreground pass of O.render_rays with a per-ray last_delta (rendering.py:199-274)."""
    n = rays.shape[0]
    rays_o, rays_d = rays[:, 0:3], rays[:, 3:6]
    near, far = rays[:, 6:7], rays[:, 7:8]
    has = last_delta.squeeze(-1) < 1e10

    def adj(zv):
        diff = torch.zeros_like(last_delta)
        diff[has, 0] = zv[has].max(dim=-1)[0]
        return last_delta - diff
    z_steps = torch.linspace(0, 1, cs)
    z = (near * (1 - z_steps) + far * z_steps).expand(n, cs)
    xyz = rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * z.unsqueeze(-1)
    out_c, l_c, _ = O._run_model_chunks(xyz, rays_d, idx, sd, cfg, "fp32", chunk, "cuda")
    comp_c = O.composite(z, out_c[..., :3], out_c[..., 3], adj(z))
    res = {"gate_loss_coarse": l_c}
    if fs == 0:
        res.update({k: comp_c[k] for k in ("rgb", "depth", "depth_variance", "bg_lambda")})
        return res
    z_mid = 0.5 * (z[:, :-1] + z[:, 1:])
    z_f = O.sample_pdf(z_mid, comp_c["weights"][:, 1:-1], fs, det=True)
    xyz_f = rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * z_f.unsqueeze(-1)
    out_f, l_f, _ = O._run_model_chunks(xyz_f, rays_d, idx, sd, cfg, "fp32", chunk, "cuda")
    z_all, order = torch.sort(torch.cat([z_f, z], -1), -1)
    rgbs = torch.gather(torch.cat([out_f[..., :3], out_c[..., :3]], 1), 1, order.unsqueeze(-1).expand(-1, -1, 3))
    sig = torch.gather(torch.cat([out_f[..., 3], out_c[..., 3]], 1), 1, order)
    comp = O.composite(z_all, rgbs, sig, adj(z_f))
    res["gate_loss_fine"] = l_f
    res.update({k: comp[k] for k in ("rgb", "depth", "depth_variance", "bg_lambda")})
    return res
