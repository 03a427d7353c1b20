"""Deterministic synthetic inputs and random-init weights of the Building topology:
seeded ray batches and state_dicts with the reference's keys/shapes.  Used by bench.py, smoke() and
the tests; there are no datasets or checkpoints offline."""
import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def synthetic_rays(n_rays: int, appearance_count: int, seed: int = 0) -> Tuple[Tensor, Tensor]:
    """o ~ U(-0.2,0.2)^3, d = normalised N(0,I), near=0.05, far=1.0."""
    g = torch.Generator().manual_seed(seed)
    o = (torch.rand(n_rays, 3, generator=g) - 0.5) * 0.4
    d = F.normalize(torch.randn(n_rays, 3, generator=g), dim=-1)
    rays = torch.cat([o, d, torch.full((n_rays, 1), 0.05), torch.full((n_rays, 1), 1.0)], 1)
    idx = torch.randint(0, appearance_count, (n_rays,), generator=g)
    return rays, idx


def synthetic_state_dict(num_experts=8, width=256, expert_layers=7, appearance_count=2048,
                         appearance_dim=48, hidden2=128, xyz_in=75, dir_in=27, seed=0,
                         gate_scale: float = 1.0) -> Dict[str, Tensor]:
    """Random-init weights with the reference's state_dict keys/shapes and
    nn.Linear-style U(-1/sqrt(in), 1/sqrt(in)) scaling.  `gate_scale` multiplies wg
    (logit temperature, config 5)."""
    g = torch.Generator().manual_seed(seed)

    def lin(out_f, in_f):
        k = 1.0 / math.sqrt(in_f)
        return ((torch.rand(out_f, in_f, generator=g) * 2 - 1) * k, (torch.rand(out_f, generator=g) * 2 - 1) * k)

    sd = {}
    for j in range(expert_layers):
        k = 1.0 / math.sqrt(width)
        sd[f"layers.0.experts.0.weights.{j}"] = (torch.rand(num_experts, width, width, generator=g) * 2 - 1) * k
        sd[f"layers.0.experts.0.bias.{j}"] = (torch.rand(num_experts, 1, width, generator=g) * 2 - 1) * k
    sd["layers.0.gates.0.wg.weight"] = lin(num_experts, width)[0] * gate_scale
    for name, (o, i) in {"layers.1": (width, width), "layers.2": (hidden2, width + dir_in + appearance_dim),
                         "layers.xyz": (width, xyz_in), "layers.sigma": (1, width), "layers.color": (3, hidden2)}.items():
        sd[f"{name}.fcs.0.weight"], sd[f"{name}.fcs.0.bias"] = lin(o, i)
    for i in range(2):
        sd[f"layers.moe_external_gate.fcs.{i}.weight"], sd[f"layers.moe_external_gate.fcs.{i}.bias"] = lin(width, width)
    sd["layers.gate_input_norm.weight"] = torch.ones(width) + 0.1 * torch.randn(width, generator=g)
    sd["layers.gate_input_norm.bias"] = 0.1 * torch.randn(width, generator=g)
    sd["embedding_a.weight"] = torch.randn(appearance_count, appearance_dim, generator=g)
    return sd


def _pe(x: Tensor, num_freqs: int) -> Tensor:
    out = [x]
    for k in range(num_freqs):
        out += [torch.sin((2.0 ** k) * x), torch.cos((2.0 ** k) * x)]
    return torch.cat(out, -1)


def _pe_mip(mu: Tensor, var: Tensor, num_freqs: int) -> Tensor:
    out = [mu]
    for k in range(num_freqs):
        damp = torch.exp(-0.5 * (4.0 ** k) * var)
        out += [torch.sin(mu * (2.0 ** k)) * damp, torch.cos(mu * (2.0 ** k)) * damp]
    return torch.cat(out, -1)


def balance_gate(sd: Dict[str, Tensor], pts: Tensor, iters: int = 60, cov: Tensor = None) -> Dict[str, Tensor]:
    """Emulate a load-balanced (trained with the l_aux balance loss) gate on random-init weights: the LayerNorm
    bias of the gate input is shifted so that wg @ beta acts as a per-expert logit offset that equalises the
    top-1 shares on `pts` [P,3] (`cov` [P,3]: mip models, integrated positional encoding).  Everything else stays the
    seeded random init.  Deterministic (CPU fp32)."""
    with torch.no_grad():
        E = sd["layers.0.gates.0.wg.weight"].shape[0]
        pe = _pe(pts, 12) if cov is None else _pe_mip(pts, cov, 12)
        h = F.linear(pe, sd["layers.xyz.fcs.0.weight"], sd["layers.xyz.fcs.0.bias"])
        g = F.linear(F.relu(F.linear(h, sd["layers.moe_external_gate.fcs.0.weight"], sd["layers.moe_external_gate.fcs.0.bias"])),
                     sd["layers.moe_external_gate.fcs.1.weight"], sd["layers.moe_external_gate.fcs.1.bias"])
        n = F.layer_norm(g, (g.shape[1],)) * sd["layers.gate_input_norm.weight"]
        wg = sd["layers.0.gates.0.wg.weight"]
        base = n @ wg.t()                                     # [P, E] logits without the LayerNorm bias
        off = wg @ sd["layers.gate_input_norm.bias"]          # current per-expert offset
        for _ in range(iters):
            share = torch.bincount(torch.argmax(base + off, 1), minlength=E).float() / base.shape[0]
            off = off - 0.5 * torch.log(share * E + 1e-2)
        beta0 = sd["layers.gate_input_norm.bias"]
        beta = beta0 + torch.linalg.pinv(wg) @ (off - wg @ beta0)
        out = dict(sd)
        out["layers.gate_input_norm.bias"] = beta.contiguous()
        return out


def benchmark_state_dict(num_experts=8, appearance_count=2048, seed=0, n_rays=8192, coarse=257, ray_seed=100):
    """Weights of the bench / reference arms: seeded random init + a gate balanced on the coarse samples of the
    benchmark's own ray batch (a trained Switch-NeRF gate is load-balanced by its auxiliary loss)."""
    sd = synthetic_state_dict(num_experts=num_experts, appearance_count=appearance_count, seed=seed, gate_scale=4.0)
    rays, _ = synthetic_rays(n_rays, appearance_count, seed=ray_seed)
    g = torch.Generator().manual_seed(seed + 7)
    pick = torch.randperm(n_rays, generator=g)[:512]
    t = torch.linspace(0, 1, 64)
    z = rays[pick, 6:7] * (1 - t) + rays[pick, 7:8] * t
    pts = (rays[pick, None, 0:3] + rays[pick, None, 3:6] * z[..., None]).reshape(-1, 3)
    return balance_gate(sd, pts)


def mission_bay_rays(n_rays: int, appearance_count: int, seed: int = 0) -> Tuple[Tensor, Tensor, Tensor]:
    """rays as synthetic_rays with near = 0.01, far = 10 scaled into the unit cube the random-init
    network sees (near 0.01, far 1.5), pixel radii ~ U(5e-4, 2e-3).  Returns rays [N,8], radii [N,1], image indices."""
    rays, idx = synthetic_rays(n_rays, appearance_count, seed)
    rays[:, 6], rays[:, 7] = 0.01, 1.5
    g = torch.Generator().manual_seed(seed + 1)
    radii = torch.rand(n_rays, 1, generator=g) * 1.5e-3 + 5e-4
    return rays, radii, idx


def mip_cast(rays: Tensor, radii: Tensor, t: Tensor) -> Tuple[Tensor, Tensor]:
    """Conical-frustum mean / diagonal covariance of the intervals between the edges t [N, S] (rendering_mip.py:15-72,
    mip-NeRF eq. 7-8 with the stable formulation); enough of it to balance the benchmark gate."""
    o, d = rays[:, None, 0:3], rays[:, None, 3:6]
    t0, t1 = t[:, :-1], t[:, 1:]
    mu, hw = (t0 + t1) / 2, (t1 - t0) / 2
    t_mean = mu + (2 * mu * hw ** 2) / (3 * mu ** 2 + hw ** 2)
    t_var = (hw ** 2) / 3 - (4 / 15) * ((hw ** 4 * (12 * mu ** 2 - hw ** 2)) / (3 * mu ** 2 + hw ** 2) ** 2)
    r_var = radii ** 2 * ((mu ** 2) / 4 + (5 / 12) * hw ** 2 - 4 / 15 * (hw ** 4) / (3 * mu ** 2 + hw ** 2))
    mean = o + d * t_mean[..., None]
    d_mag = (d ** 2).sum(-1, keepdim=True).clamp_min(1e-10)
    cov = t_var[..., None] * d ** 2 + r_var[..., None] * (1 - d ** 2 / d_mag)
    return mean, cov


def mission_bay_state_dict(num_experts=8, appearance_count=2048, seed=0, n_rays=13312, coarse=257, ray_seed=100):
    """Mission-Bay topology (mission_bay.yaml: width 512, MipNeRFMoE) with a gate balanced on the benchmark's rays."""
    sd = synthetic_state_dict(num_experts=num_experts, width=512, appearance_count=appearance_count, seed=seed, gate_scale=4.0)
    rays, radii, _ = mission_bay_rays(n_rays, appearance_count, seed=ray_seed)
    g = torch.Generator().manual_seed(seed + 7)
    pick = torch.randperm(n_rays, generator=g)[:512]
    t = torch.linspace(0, 1, 65)
    z = rays[pick, 6:7] * (1 - t) + rays[pick, 7:8] * t
    mean, cov = mip_cast(rays[pick], radii[pick], z)
    return balance_gate(sd, mean.reshape(-1, 3), cov=cov.reshape(-1, 3))
