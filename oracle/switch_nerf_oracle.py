"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch-CPU tensors, no CUDA) of the
Switch-NeRF forward/render hot path.  Nothing under `switch_nerf_b200/` imports
this file; it is the checker for `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs.

Parity status: PINNED.  Every function below is checked against the unmodified
reference (`/root/reference/switch_nerf`, imported through `oracle/ref_shims.py`)
by `tests/test_oracle_vs_reference.py` (runs when /root/reference exists) and
against the committed fixtures `tests/golden/*.npz` written by
`oracle/make_golden.py` from that same unmodified reference.  The one boundary
that cannot be pinned is the external Tutel binary (microsoft/tutel @ 56dbd664,
not vendored): its dispatch/combine kernels are restated from the reference's
call sites and from the in-tree no-batch kernels with identical addressing
(tutel_sparse_nobatch.py:24-34, 45-63).

All `file:line` citations are relative to /root/reference/switch_nerf/.

Precision modes
---------------
mode="fp32"  : everything in fp32 -- what the reference does on CPU.
mode="bf16"  : the op-by-op precision map of `torch.autocast(bf16)` around the
               reference (SURVEY.md §8a footnote): Linear/baddbmm take bf16
               operands, accumulate in fp32 and round the result to bf16;
               LayerNorm, gate GEMM, softmax, softplus and compositing are fp32.
               `flavor="cpu"` reproduces torch's *CPU* autocast op lists (LayerNorm
               and softplus stay in bf16) so that this mode can be pinned against
               the reference run under `torch.autocast("cpu", bfloat16)` here.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- #
# encoders / activations                                                       #
# --------------------------------------------------------------------------- #
def embedding(x: Tensor, num_freqs: int) -> Tensor:
    """models/nerf.py:9-26  Embedding.forward: [x, sin(2^k x), cos(2^k x)]_k<num_freqs."""
    out = [x]
    for k in range(num_freqs):
        f = float(2 ** k)
        out += [torch.sin(f * x), torch.cos(f * x)]
    return torch.cat(out, -1)


def mip_embedding(x: Tensor, num_freqs: int, d: int = 3) -> Tensor:
    """models/nerf.py:28-56  MipEmbedder: sin/cos(2^k mu) * exp(-0.5 * 4^k * var)."""
    mu, var = x[:, :d], x[:, d:]
    out = [mu]
    for k in range(num_freqs):
        fy, fw = float(2.0 ** k), float(4.0 ** k)
        damp = torch.exp(-0.5 * fw * var)
        out += [torch.sin(mu * fy) * damp, torch.cos(mu * fy) * damp]
    return torch.cat(out, -1)


def shifted_softplus(x: Tensor) -> Tensor:
    """models/nerf.py:58-72: softplus(x - 1), beta=1, threshold=20."""
    return F.softplus(x - 1, 1, 20)


# --------------------------------------------------------------------------- #
# precision helpers                                                            #
# --------------------------------------------------------------------------- #
def _r(x: Tensor) -> Tensor:
    """Round to bf16 and come back to fp32 (value-preserving container)."""
    return x.to(torch.bfloat16).to(torch.float32)


def linear(x: Tensor, w: Tensor, b: Optional[Tensor], mode: str) -> Tensor:
    """nn.Linear with weight [out,in].  bf16 mode = autocast semantics: operands
    rounded to bf16, fp32 accumulation, result rounded to bf16."""
    if mode == "fp32":
        return F.linear(x, w, b)
    y = F.linear(_r(x), _r(w), None if b is None else _r(b))
    return _r(y)


def mlp(x: Tensor, sd: Dict[str, Tensor], prefix: str, num: int, mode: str) -> Tensor:
    """models/nerf_moe.py:30-49  Mlp.forward without skips (none of the configs'
    dense Mlps has skips): ReLU between layers, none after the last."""
    h = x
    for i in range(num):
        h = linear(h, sd[f"{prefix}.fcs.{i}.weight"], sd[f"{prefix}.fcs.{i}.bias"], mode)
        if i < num - 1:
            h = F.relu(h)
    return h


# --------------------------------------------------------------------------- #
# routing (Appendix A of SURVEY.md)                                            #
# --------------------------------------------------------------------------- #
def capacity_of(num_samples: int, num_experts: int, capacity_factor: float, top_k: int = 1) -> int:
    """tutel_fast_dispatch.py:210-211."""
    return top_k * int(capacity_factor * ((int(num_samples) + num_experts - 1) // num_experts))


def route_top1(gates: Tensor, capacity_factor: float, bpr: bool):
    """tutel_fast_dispatch.py:176-217 extract_critical for k=1 (+131-150 helpers).

    Returns idx int32[S], loc int32[S], gate fp32[S], capacity int, l_aux fp32 scalar.
    Tie-break contract: argmax -> lowest expert
    index; BPR order -> descending max-gate, ties by ascending sample index.
    """
    S, E = gates.shape
    idx = torch.argmax(gates, dim=1)                                    # :177 (topk k=1)
    mask = F.one_hot(idx, E).to(torch.int64)                            # :131-134
    gate_val = (gates * mask).sum(dim=1)                                # :182
    me = torch.sum(gates.float(), dim=0)                                # :141-145 (fp32 branch)
    ce = torch.sum(mask.to(me.dtype), dim=0)
    l_aux = torch.sum(me * ce) * (E / (S * S))
    if bpr:                                                             # :186-188, 136-139
        importance = -1 * gates.max(dim=1)[0]
        order = importance.argsort(dim=0, stable=True)
        sorted_mask = mask[order]
        sorted_cumsum = (torch.cumsum(sorted_mask, 0) - 1) * sorted_mask
        loc1 = sorted_cumsum[order.argsort(dim=0, stable=True)]
    else:                                                               # :190
        loc1 = torch.cumsum(mask, 0) - 1
    loc = torch.sum(loc1 * mask, dim=1).to(torch.int32)                 # :194
    cap = capacity_of(S, E, capacity_factor)
    return idx.to(torch.int32), loc, gate_val, cap, l_aux


def route_top1_nobatch(gates: Tensor):
    """tutel_fast_dispatch_nobatch.py:205-251 for k=1, no BPR: additionally the
    per-expert counts (expert_input_nums, :225) and their exclusive cumsum (:26-29)."""
    S, E = gates.shape
    idx, loc, gate_val, _, l_aux = route_top1(gates, 1.0, False)
    counts = torch.bincount(idx.long(), minlength=E).to(torch.int32)
    begin = (torch.cumsum(counts, 0) - counts).to(torch.int32)
    return idx, loc, gate_val, counts, begin, l_aux


def dispatch(x: Tensor, idx: Tensor, loc: Tensor, num_experts: int, cap: int) -> Tensor:
    """GatingEncoder.forward (tutel_fast_dispatch.py:15-28) with is_postscore=True:
    zeros[E*cap, M]; row idx*cap+loc <- x[s] iff loc < cap."""
    buf = torch.zeros(num_experts * cap, x.shape[1], dtype=x.dtype)
    keep = loc < cap
    rows = idx.long() * cap + loc.long()
    buf[rows[keep]] = x[keep]
    return buf


def combine(buf: Tensor, idx: Tensor, loc: Tensor, gate_val: Tensor, cap: int) -> Tensor:
    """GatingDecoder.forward (tutel_fast_dispatch.py:48-63): y[s] = gate*buf[row], dropped -> 0."""
    keep = loc < cap
    rows = (idx.long() * cap + loc.long()).clamp(0, buf.shape[0] - 1)
    y = buf[rows] * gate_val.unsqueeze(1).to(buf.dtype)
    return torch.where(keep.unsqueeze(1), y, torch.zeros_like(y))


def expert_mlp(h: Tensor, weights, biases, skips, mode: str) -> Tensor:
    """ExpertMLP.forward (tutel_moe_layer_nobatch.py:887-924).
    h [E, cap, M]; weights[j] [E, M(in), M(out)]; biases[j] [E, 1, M]."""
    x = h
    n = len(weights)
    for j in range(n):
        w, b = weights[j], biases[j]
        if mode == "fp32":
            h = torch.baddbmm(b, h, w)
        else:
            h = _r(torch.baddbmm(_r(b), _r(h), _r(w)))
        if skips is not None and j in skips:
            h = h + x
            if mode != "fp32":
                h = _r(h)
            if j < n - 1:
                h = F.relu(h)
            x = h
        elif j < n - 1:
            h = F.relu(h)
    return h


def moe_layer(h: Tensor, gate_input: Tensor, sd: Dict[str, Tensor], tag: str, cfg: dict, mode: str):
    """MOELayer.forward -> TopKGate.apply_on_expert_fn / _nobatch
    (tutel_moe_layer_nobatch.py:733-797, 98-235, 237-352)."""
    E = cfg["num_experts"]
    wg = sd[f"layers.{tag}.gates.0.wg.weight"].float()
    logits = F.linear(gate_input.float(), wg)                            # :105-113 fp32, autocast off
    gates = F.softmax(logits, dim=1)                                     # :126
    weights = [sd[f"layers.{tag}.experts.0.weights.{j}"] for j in range(cfg["expert_layers"])]
    biases = [sd[f"layers.{tag}.experts.0.bias.{j}"] for j in range(cfg["expert_layers"])]
    S, M = h.shape
    if cfg.get("moe_no_batch", False):
        idx, loc, gate_val, counts, begin, l_aux = route_top1_nobatch(gates)
        cap = int(counts.max().item()) if S > 0 else 0
        # contiguous-by-expert buffer == padded buffer with cap=max count, nothing dropped
        buf = dispatch(h.float(), idx, loc, E, max(cap, 1))
        cap = max(cap, 1)
    else:
        idx, loc, gate_val, cap, l_aux = route_top1(gates, cfg["capacity_factor"], cfg["bpr"])
        buf = dispatch(h.float(), idx, loc, E, cap)                       # dispatch dtype fp32 (:89-92)
    if mode != "fp32":
        buf = _r(buf)                                                     # .to(original_dtype) (:119)
    out = expert_mlp(buf.view(E, cap, M), weights, biases, cfg["skips"], mode).reshape(E * cap, M)
    y = combine(out.float(), idx, loc, gate_val.float(), cap)            # fp32 * fp32 gate (:123-127)
    if mode != "fp32":
        y = _r(y)
    extras = {"gates": gates, "logits": logits, "idx": idx, "loc": loc, "gate_val": gate_val,
              "capacity": cap, "l_aux": l_aux}
    return y, extras


# --------------------------------------------------------------------------- #
# model forward                                                                #
# --------------------------------------------------------------------------- #
def default_cfg(sd: Dict[str, Tensor], capacity_factor=1.0, bpr=True, moe_no_batch=False,
                mip=False) -> dict:
    """Derive the topology constants from a reference state_dict."""
    E = sd["layers.0.gates.0.wg.weight"].shape[0]
    n_exp = len([k for k in sd if k.startswith("layers.0.experts.0.weights.")])
    return {"num_experts": E, "expert_layers": n_exp, "skips": [3], "capacity_factor": capacity_factor,
            "bpr": bpr, "moe_no_batch": moe_no_batch, "pos_xyz_dim": 12, "pos_dir_dim": 4, "mip": mip,
            "gate_layers": len([k for k in sd if k.startswith("layers.moe_external_gate.fcs.") and k.endswith("weight")])}


def nerf_moe_forward(x: Tensor, sd: Dict[str, Tensor], cfg: dict, mode: str = "fp32",
                     sigma_noise: Optional[Tensor] = None, flavor: str = "cuda"):
    """NeRFMoE.forward / MipNeRFMoE.forward (models/nerf_moe.py:320-455 / 675-810),
    Building/Mission-Bay topology.  x = [xyz(3) (+cov(3) for mip), dir(3), image_index(1)].
    Returns outputs [S,4] fp32 and the routing extras of the single MoE layer."""
    xd = 6 if cfg.get("mip") else 3
    if x.shape[1] != xd + 4:
        raise Exception("Unexpected input shape: {} (expected: {}, xyz_dim: {})".format(x.shape, xd + 4, xd))
    pe = mip_embedding(x[:, :xd], cfg["pos_xyz_dim"]) if cfg.get("mip") else embedding(x[:, :3], cfg["pos_xyz_dim"])
    h = mlp(pe, sd, "layers.xyz", 1, mode)                                # :330-333 (act none)
    g = mlp(h, sd, "layers.moe_external_gate", cfg["gate_layers"], mode)  # :346-348
    lw, lb = sd["layers.gate_input_norm.weight"], sd["layers.gate_input_norm.bias"]
    if mode == "bf16" and flavor == "cpu":
        gate_input = _r(F.layer_norm(g, (g.shape[1],), _r(lw), _r(lb)))   # CPU autocast: LN stays bf16
    else:
        gate_input = F.layer_norm(g.float(), (g.shape[1],), lw.float(), lb.float())  # :372 (fp32 under cuda autocast)
    h, ex = moe_layer(h, gate_input, sd, "0", cfg, mode)                  # :374
    h = F.relu(h)                                                         # :384-386
    sigma = linear(h, sd["layers.sigma.fcs.0.weight"], sd["layers.sigma.fcs.0.bias"], mode)  # :396-397
    if sigma_noise is not None:
        sigma = sigma + sigma_noise                                       # in-place on the (bf16) Linear output, :407-408
        if mode == "bf16":
            sigma = _r(sigma)
    if mode == "bf16" and flavor == "cpu":
        sigma = _r(shifted_softplus(_r(sigma)))
    elif mode == "bf16":
        # cuda autocast: `sigma += noise` and `x - 1` (nerf.py:68) are plain bf16 tensor ops -- each rounds to bf16 --
        # and only F.softplus itself is promoted to fp32.  Pinned by tests/golden/model_*_bf16cuda.npz (the unmodified
        # reference on a B200): without the rounding of x - 1, 97 % of the sigmas are off by up to 1e-3.
        sigma = F.softplus(_r(_r(sigma) - 1.0).float())
    else:
        sigma = shifted_softplus(sigma.float())                           # :416
    h = mlp(h, sd, "layers.1", 1, mode)                                   # layer "1", act none
    d_pe = embedding(x[:, xd:xd + 3], cfg["pos_dir_dim"])                 # :424
    emb = sd["embedding_a.weight"][x[:, -1].long()]                       # :427
    h = torch.cat([h, d_pe, emb], -1)                                     # :429
    h = F.relu(mlp(h, sd, "layers.2", 1, mode))                           # layer "2", act relu
    rgb = linear(h, sd["layers.color.fcs.0.weight"], sd["layers.color.fcs.0.bias"], mode)    # :434
    rgb = torch.sigmoid(rgb)
    if mode != "fp32":
        rgb = _r(rgb)
    outputs = torch.cat([rgb.float(), sigma.float()], -1)                 # :441
    return outputs, ex


# --------------------------------------------------------------------------- #
# rendering                                                                    #
# --------------------------------------------------------------------------- #
def expand_and_perturb_z_vals(z_vals: Tensor, samples: int, perturb: float, n_rays: int,
                              rand: Optional[Tensor] = None) -> Tensor:
    """rendering.py:573-584.  `rand` replaces torch.rand_like for reproducibility."""
    z_vals = z_vals.expand(n_rays, samples)
    if perturb > 0:
        mid = 0.5 * (z_vals[:, :-1] + z_vals[:, 1:])
        upper = torch.cat([mid, z_vals[:, -1:]], -1)
        lower = torch.cat([z_vals[:, :1], mid], -1)
        r = torch.rand_like(z_vals) if rand is None else rand
        z_vals = lower + (upper - lower) * (perturb * r)
    return z_vals


def sample_pdf(bins: Tensor, weights: Tensor, fine_samples: int, det: bool,
               u: Optional[Tensor] = None) -> Tensor:
    """rendering.py:587-637 (_sample_pdf + _sample_cdf)."""
    weights = weights + 1e-8
    pdf = weights / weights.sum(-1).unsqueeze(-1)
    cdf = torch.cumsum(pdf, -1)
    n_rays, n_s = cdf.shape
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)
    if det:
        u = torch.linspace(0, 1, fine_samples).expand(n_rays, fine_samples)
    elif u is None:
        u = torch.rand(n_rays, fine_samples)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp_min(inds - 1, 0)
    above = torch.clamp_max(inds, n_s)
    cdf0, cdf1 = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    b0, b1 = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf1 - cdf0
    denom = torch.where(denom < 1e-8, torch.ones_like(denom), denom)
    return b0 + (u - cdf0) / denom * (b1 - b0)


def composite(z_vals: Tensor, rgbs: Tensor, sigmas: Tensor, last_delta: Tensor):
    """rendering.py:436-494 (flip=False): alpha, exclusive transmittance, weights,
    rgb / depth / depth variance, bg_lambda."""
    deltas = torch.cat([z_vals[:, 1:] - z_vals[:, :-1], last_delta], -1)
    alphas = 1 - torch.exp(-deltas * sigmas)
    T = torch.cumprod(1 - alphas + 1e-8, -1)
    bg_lambda = T[..., -1]
    T = torch.cat((torch.ones_like(T[..., 0:1]), T[..., :-1]), dim=-1)
    weights = alphas * T
    rgb = (weights.unsqueeze(-1) * rgbs).sum(dim=1)
    depth = (weights * z_vals).sum(dim=1)
    var = (weights * (z_vals - depth.unsqueeze(1)).square()).sum(-1)
    return {"weights": weights, "rgb": rgb, "depth": depth, "depth_variance": var,
            "bg_lambda": bg_lambda, "alphas": alphas}


def _run_model_chunks(xyz: Tensor, rays_d: Tensor, image_indices: Tensor, sd, cfg, mode, chunk, flavor):
    """rendering.py:306-314, 354-409: flatten, repeat dirs / indices, chunk loop."""
    n_rays, n_s = xyz.shape[:2]
    x = torch.cat([xyz.reshape(-1, xyz.shape[-1]),
                   rays_d.view(n_rays, 1, 3).expand(n_rays, n_s, 3).reshape(-1, 3),
                   image_indices.view(n_rays, 1, 1).expand(n_rays, n_s, 1).reshape(-1, 1).to(xyz.dtype)], 1)
    outs, l_aux, idxs = [], [], []
    for i in range(0, x.shape[0], chunk):
        o, ex = nerf_moe_forward(x[i:i + chunk], sd, cfg, mode, flavor=flavor)
        outs.append(o)
        l_aux.append(ex["l_aux"].reshape(1))
        idxs.append(ex["idx"].long().view(-1, 1, 1))
    out = torch.cat(outs, 0).view(n_rays, n_s, 4)
    return out, torch.cat(l_aux, 0), torch.cat(idxs, 0).view(n_rays, n_s, 1, 1)


def render_rays(sd: Dict[str, Tensor], cfg: dict, rays: Tensor, image_indices: Tensor, *,
                coarse_samples: int, fine_samples: int, model_chunk_size: int, mode: str = "fp32",
                perturb: float = 0.0, rand_coarse: Optional[Tensor] = None, u_fine: Optional[Tensor] = None,
                flavor: str = "cuda") -> Dict[str, Tensor]:
    """rendering.render_rays (rendering.py:15-196) for bg_nerf=None, use_cascade=False,
    plus _get_results (199-274) and _inference (277-494)."""
    n_rays = rays.shape[0]
    rays_o, rays_d = rays[:, 0:3], rays[:, 3:6]
    near, far = rays[:, 6:7], rays[:, 7:8]
    last_delta = 1e10 * torch.ones(n_rays, 1)
    z_steps = torch.linspace(0, 1, coarse_samples)                                       # :85
    z_vals = near * (1 - z_steps) + far * z_steps                                        # :86
    z_vals = expand_and_perturb_z_vals(z_vals, coarse_samples, perturb, n_rays, rand_coarse)
    xyz = rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * z_vals.unsqueeze(-1)                # :90
    res = {}
    out_c, l_aux_c, gates_c = _run_model_chunks(xyz, rays_d, image_indices, sd, cfg, mode, model_chunk_size, flavor)
    res["gate_loss_coarse"], res["moe_gates_coarse"] = l_aux_c, gates_c
    comp_c = composite(z_vals, out_c[..., :3], out_c[..., 3], last_delta)
    res["_z_coarse"], res["_raw_coarse"], res["_weights_coarse"] = z_vals, out_c, comp_c["weights"]
    if fine_samples == 0:
        for k in ("rgb", "depth", "depth_variance"):
            res[f"{k}_coarse"] = comp_c[k]
        return res
    z_mid = 0.5 * (z_vals[:, :-1] + z_vals[:, 1:])                                        # :238
    # rendering.py:240 samples from the DETACHED coarse weights: no gradient flows through the sample positions
    z_fine = sample_pdf(z_mid, comp_c["weights"][:, 1:-1].detach(), fine_samples, det=(perturb == 0), u=u_fine)
    xyz_f = rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * z_fine.unsqueeze(-1)
    out_f, l_aux_f, gates_f = _run_model_chunks(xyz_f, rays_d, image_indices, sd, cfg, mode, model_chunk_size, flavor)
    res["gate_loss_fine"], res["moe_gates_fine"] = l_aux_f, gates_f
    z_all, order = torch.sort(torch.cat([z_fine, z_vals], -1), -1)                        # :421
    rgbs = torch.gather(torch.cat([out_f[..., :3], out_c[..., :3]], 1), 1, order.unsqueeze(-1).expand(-1, -1, 3))
    sig = torch.gather(torch.cat([out_f[..., 3], out_c[..., 3]], 1), 1, order)
    comp = composite(z_all, rgbs, sig, last_delta)
    for k in ("rgb", "depth", "depth_variance"):
        res[f"{k}_fine"] = comp[k]
    res["_z_fine"], res["_raw_fine"] = z_fine, out_f
    return res



# --------------------------------------------------------------------------- #
# mip renderer (rendering_mip.py) -- Mission Bay / Bungee configs              #
# --------------------------------------------------------------------------- #
def mip_cast_rays(origin: Tensor, direction: Tensor, radius: Tensor, t: Tensor):
    """rendering_mip.py:15-25: conical-frustum mean / diagonal covariance per interval."""
    t0, t1 = t[..., :-1], t[..., 1:]
    c, d = (t0 + t1) / 2, (t1 - t0) / 2
    t_mean = c + (2 * c * d ** 2) / (3 * c ** 2 + d ** 2)
    t_var = (d ** 2) / 3 - (4 / 15) * ((d ** 4 * (12 * c ** 2 - d ** 2)) / (3 * c ** 2 + d ** 2) ** 2)
    r_var = radius ** 2 * ((c ** 2) / 4 + (5 / 12) * d ** 2 - (4 / 15) * (d ** 4) / (3 * c ** 2 + d ** 2))
    mean = origin[..., None, :] + direction[..., None, :] * t_mean[..., None]
    null_outer_diag = 1 - (direction ** 2) / torch.sum(direction ** 2, -1, keepdims=True)
    cov_diag = t_var[..., None] * (direction ** 2)[..., None, :] + r_var[..., None] * null_outer_diag[..., None, :]
    return mean, cov_diag


def sorted_piecewise_constant_pdf(bins: Tensor, weights: Tensor, num_samples: int, u: Optional[Tensor] = None) -> Tensor:
    """rendering_mip.py:75-131 (sorted_piecewise_constant_pdf1), deterministic branch unless `u` is given.
    The O(N*S^2) boolean mask of the reference is restated with searchsorted (same selected edges)."""
    eps = 1e-5
    weights = weights.clone()
    weight_sum = torch.sum(weights, dim=-1, keepdim=True)
    padding = torch.clamp_min(eps - weight_sum, 0)
    weights = weights + padding / weights.shape[-1]
    weight_sum = weight_sum + padding
    pdf = weights / weight_sum
    cdf = torch.clamp_max(torch.cumsum(pdf[..., :-1], dim=-1), 1.0)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf, torch.ones_like(cdf[..., :1])], -1)
    if u is None:
        u = torch.linspace(0., 1. - torch.finfo(torch.float32).eps, num_samples)
        u = u.expand(list(cdf.shape[:-1]) + [num_samples])
    u = u.contiguous()
    n = cdf.shape[-1]
    hi = torch.searchsorted(cdf, u, right=True)          # first index with cdf > u
    i0 = torch.clamp(hi - 1, 0, n - 1)                   # last index with cdf <= u (cdf[0] = 0 <= u always)
    i1 = torch.clamp(hi, 0, n - 1)                       # no such index -> last edge (x[..., -1:])
    b0, b1 = torch.gather(bins, -1, i0), torch.gather(bins, -1, i1)
    c0, c1 = torch.gather(cdf, -1, i0), torch.gather(cdf, -1, i1)
    t = torch.clip(torch.nan_to_num((u - c0) / (c1 - c0), 0), 0, 1)
    return b0 + t * (b1 - b0)


def mip_composite(z_edges: Tensor, rgbs: Tensor, sigmas: Tensor, last_delta: Tensor, rgb_padding: Optional[float]):
    """rendering_mip.py:382-425: composite on interval mid points with rgb padding."""
    if rgb_padding is not None:
        rgbs = rgbs * (1 + 2 * rgb_padding) - rgb_padding
    z = .5 * (z_edges[..., 1:] + z_edges[..., :-1])
    return composite(z, rgbs, sigmas, last_delta)


def _run_mip_chunks(mean, cov, rays_d, image_indices, sd, cfg, mode, chunk, flavor):
    n_rays, n_s = mean.shape[:2]
    x = torch.cat([mean.reshape(-1, 3), cov.reshape(-1, 3),
                   rays_d.view(n_rays, 1, 3).expand(n_rays, n_s, 3).reshape(-1, 3),
                   image_indices.view(n_rays, 1, 1).expand(n_rays, n_s, 1).reshape(-1, 1).to(mean.dtype)], 1)
    outs, l_aux, idxs = [], [], []
    for i in range(0, x.shape[0], chunk):
        o, ex = nerf_moe_forward(x[i:i + chunk], sd, cfg, mode, flavor=flavor)
        outs.append(o)
        l_aux.append(ex["l_aux"].reshape(1))
        idxs.append(ex["idx"].long().view(-1, 1, 1))
    return torch.cat(outs, 0).view(n_rays, n_s, 4), torch.cat(l_aux, 0), torch.cat(idxs, 0).view(n_rays, n_s, 1, 1)


def render_rays_mip(sd: Dict[str, Tensor], cfg: dict, rays: Tensor, radii: Tensor, image_indices: Tensor, *,
                    coarse_samples: int, fine_samples: int, model_chunk_size: int, mode: str = "fp32",
                    weights_resample_padding: float = 0.01, rgb_padding: Optional[float] = 0.001,
                    flavor: str = "cuda", stop_level_grad: bool = True) -> Dict[str, Tensor]:
    """rendering_mip.render_rays (133-174) + _get_results (177-261) + _inference (264-425), eval mode
    (perturb = 0, deterministic resampling)."""
    n_rays = rays.shape[0]
    rays_o, rays_d = rays[:, 0:3], rays[:, 3:6]
    near, far = rays[:, 6:7], rays[:, 7:8]
    last_delta = 1e10 * torch.ones(n_rays, 1)
    z_steps = torch.linspace(0, 1, coarse_samples)
    z_vals = (near * (1 - z_steps) + far * z_steps).expand(n_rays, coarse_samples)
    res = {}
    mean, cov = mip_cast_rays(rays_o, rays_d, radii, z_vals)
    out_c, l_c, g_c = _run_mip_chunks(mean, cov, rays_d, image_indices, sd, cfg, mode, model_chunk_size, flavor)
    res["gate_loss_coarse"], res["moe_gates_coarse"] = l_c, g_c
    comp_c = mip_composite(z_vals, out_c[..., :3], out_c[..., 3], last_delta, rgb_padding)
    res["rgb_coarse"] = comp_c["rgb"]
    res["_raw_coarse"] = out_c
    if fine_samples == 0:
        res["depth_coarse"], res["depth_variance_coarse"] = comp_c["depth"], comp_c["depth_variance"]
        return res
    w = comp_c["weights"]
    w_pad = torch.cat([w[..., :1], w, w[..., -1:]], -1)                       # :218-222
    w_max = torch.maximum(w_pad[..., :-1], w_pad[..., 1:])
    w_blur = 0.5 * (w_max[..., :-1] + w_max[..., 1:])
    z_samples = sorted_piecewise_constant_pdf(z_vals, w_blur + weights_resample_padding, fine_samples)
    if stop_level_grad:                                                        # :227-228 (opts.py:245: always true)
        z_samples = z_samples.detach()
    z_fine, _ = torch.sort(z_samples, -1)
    mean, cov = mip_cast_rays(rays_o, rays_d, radii, z_fine)
    out_f, l_f, g_f = _run_mip_chunks(mean, cov, rays_d, image_indices, sd, cfg, mode, model_chunk_size, flavor)
    res["gate_loss_fine"], res["moe_gates_fine"] = l_f, g_f
    comp = mip_composite(z_fine, out_f[..., :3], out_f[..., 3], last_delta, rgb_padding)
    res["rgb_fine"], res["depth_fine"], res["depth_variance_fine"] = comp["rgb"], comp["depth"], comp["depth_variance"]
    res["_z_fine"], res["_raw_fine"] = z_fine, out_f
    return res

def psnr(a: Tensor, b: Tensor) -> float:
    """metrics.py:8-10."""
    return float(-10.0 * torch.log10(torch.mean((a - b) ** 2)))


# deterministic synthetic inputs / weights live in the product package (bench.py's GPU arm uses them
# too and must not import anything under oracle/); re-exported here for the tests.
from switch_nerf_b200.synthetic import synthetic_rays, synthetic_state_dict  # noqa: E402,F401
