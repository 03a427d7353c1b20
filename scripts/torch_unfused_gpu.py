"""GPU comparator (SURVEY §8d "reference PyTorch + Tutel-shim on B200"): the same render step written the way the
reference runs it -- one PyTorch op per stage, bf16 autocast, cuBLAS GEMMs, [E*cap, M] dispatch buffer,
sort + gather merge -- timed on the same B200 as bench.py.  Real Tutel cannot be built offline, so its two sparse
kernels are index_put / gather here (as in SURVEY §8c's shim); every other op is what upstream launches.

Self-contained on purpose (no import from oracle/): it is a measurement aid, not the parity oracle, and not part
of the product.  It also reports max |rgb - ours| on the same rays so the two arms are known to compute the same
image.   python scripts/torch_unfused_gpu.py [--steps 5 --warmup 2 --rays 8192]  -> one JSON line."""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

AC = dict(device_type="cuda", dtype=torch.bfloat16)
ACT = torch.bfloat16      # dtype of the dispatched / combined activations (autocast's .to(original_dtype))


def pe(x, n):
    out = [x]
    for k in range(n):
        out += [torch.sin((2.0 ** k) * x), torch.cos((2.0 ** k) * x)]
    return torch.cat(out, -1)


def route(gates, cf, bpr):
    S, E = gates.shape
    idx = torch.argmax(gates, 1)
    mask = F.one_hot(idx, E).to(torch.int64)
    gate_val = (gates * mask).sum(1)
    l_aux = torch.sum(gates.sum(0) * mask.float().sum(0)) * (E / (S * S))
    if bpr:
        order = (-gates.max(1)[0]).argsort(dim=0, stable=True)
        sm = mask[order]
        loc1 = ((torch.cumsum(sm, 0) - 1) * sm)[order.argsort(dim=0, stable=True)]
    else:
        loc1 = torch.cumsum(mask, 0) - 1
    loc = (loc1 * mask).sum(1)
    return idx, loc, gate_val, int(cf * ((S + E - 1) // E)), l_aux


def model_chunk(x, sd, cf, bpr, n_exp_layers=7, skip=3):
    E = sd["layers.0.gates.0.wg.weight"].shape[0]
    with torch.autocast(**AC):
        h = F.linear(pe(x[:, :3], 12), sd["layers.xyz.fcs.0.weight"], sd["layers.xyz.fcs.0.bias"])
        g = F.linear(h, sd["layers.moe_external_gate.fcs.0.weight"], sd["layers.moe_external_gate.fcs.0.bias"])
        g = F.linear(F.relu(g), sd["layers.moe_external_gate.fcs.1.weight"], sd["layers.moe_external_gate.fcs.1.bias"])
        gi = F.layer_norm(g, (g.shape[1],), sd["layers.gate_input_norm.weight"], sd["layers.gate_input_norm.bias"])
    gates = F.softmax(F.linear(gi.float(), sd["layers.0.gates.0.wg.weight"]), 1)
    idx, loc, gate_val, cap, l_aux = route(gates, cf, bpr)
    keep = loc < cap
    rows = idx * cap + loc
    buf = torch.zeros(E * cap, h.shape[1], dtype=torch.float32, device=x.device)
    buf[rows[keep]] = h.float()[keep]
    t = buf.to(ACT).view(E, cap, -1)
    with torch.autocast(**AC):
        xin = t
        for j in range(n_exp_layers):
            t = torch.baddbmm(sd[f"layers.0.experts.0.bias.{j}"], t, sd[f"layers.0.experts.0.weights.{j}"])
            if j == skip:
                t = t + xin
                t = F.relu(t)
                xin = t
            elif j < n_exp_layers - 1:
                t = F.relu(t)
    y = t.reshape(E * cap, -1).float()[rows.clamp(0, E * cap - 1)] * gate_val.unsqueeze(1)
    y = torch.where(keep.unsqueeze(1), y, torch.zeros_like(y)).to(ACT)
    with torch.autocast(**AC):
        hr = F.relu(y)
        sigma = F.linear(hr, sd["layers.sigma.fcs.0.weight"], sd["layers.sigma.fcs.0.bias"])
        sigma = F.softplus(sigma.float() - 1, 1, 20)
        h1 = F.linear(hr, sd["layers.1.fcs.0.weight"], sd["layers.1.fcs.0.bias"])
        cat = torch.cat([h1, pe(x[:, 3:6], 4), sd["embedding_a.weight"][x[:, -1].long()]], -1)
        h2 = F.relu(F.linear(cat, sd["layers.2.fcs.0.weight"], sd["layers.2.fcs.0.bias"]))
        rgb = torch.sigmoid(F.linear(h2, sd["layers.color.fcs.0.weight"], sd["layers.color.fcs.0.bias"]))
    return torch.cat([rgb.float(), sigma.float()], -1), l_aux


def run_chunks(xyz, rays_d, image_indices, sd, chunk, cf, bpr):
    n, s = xyz.shape[:2]
    x = torch.cat([xyz.reshape(-1, 3), rays_d.view(n, 1, 3).expand(n, s, 3).reshape(-1, 3),
                   image_indices.view(n, 1, 1).expand(n, s, 1).reshape(-1, 1).float()], 1)
    outs = [model_chunk(x[i:i + chunk], sd, cf, bpr)[0] for i in range(0, x.shape[0], chunk)]
    return torch.cat(outs, 0).view(n, s, 4)


def composite(z, rgbs, sig, last_delta):
    deltas = torch.cat([z[:, 1:] - z[:, :-1], last_delta], -1)
    alphas = 1 - torch.exp(-deltas * sig)
    T = torch.cumprod(1 - alphas + 1e-8, -1)
    T = torch.cat((torch.ones_like(T[..., 0:1]), T[..., :-1]), -1)
    w = alphas * T
    return w, (w.unsqueeze(-1) * rgbs).sum(1), (w * z).sum(1)


def sample_pdf(bins, weights, n_fine):
    weights = weights + 1e-8
    cdf = torch.cumsum(weights / weights.sum(-1, keepdim=True), -1)
    n_rays, n_s = cdf.shape
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)
    u = torch.linspace(0, 1, n_fine, device=bins.device).expand(n_rays, n_fine).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below, above = torch.clamp_min(inds - 1, 0), torch.clamp_max(inds, n_s)
    c0, c1 = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    b0, b1 = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    den = c1 - c0
    den = torch.where(den < 1e-8, torch.ones_like(den), den)
    return b0 + (u - c0) / den * (b1 - b0)


def render(sd, rays, image_indices, coarse, fine, chunk, cf=1.0, bpr=True):
    n = rays.shape[0]
    o, d, near, far = rays[:, 0:3], rays[:, 3:6], rays[:, 6:7], rays[:, 7:8]
    last_delta = 1e10 * torch.ones(n, 1, device=rays.device)
    t = torch.linspace(0, 1, coarse, device=rays.device)
    z = (near * (1 - t) + far * t).expand(n, coarse)
    out_c = run_chunks(o.unsqueeze(1) + d.unsqueeze(1) * z.unsqueeze(-1), d, image_indices, sd, chunk, cf, bpr)
    w_c, _, _ = composite(z, out_c[..., :3], out_c[..., 3], last_delta)
    z_mid = 0.5 * (z[:, :-1] + z[:, 1:])
    z_f = sample_pdf(z_mid, w_c[:, 1:-1], fine)
    out_f = run_chunks(o.unsqueeze(1) + d.unsqueeze(1) * z_f.unsqueeze(-1), d, image_indices, sd, chunk, cf, bpr)
    z_all, order = torch.sort(torch.cat([z_f, z], -1), -1)
    rgbs = torch.gather(torch.cat([out_f[..., :3], out_c[..., :3]], 1), 1, order.unsqueeze(-1).expand(-1, -1, 3))
    sig = torch.gather(torch.cat([out_f[..., 3], out_c[..., 3]], 1), 1, order)
    _, rgb, depth = composite(z_all, rgbs, sig, last_delta)
    return rgb, depth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--rays", type=int, default=8192)
    args = ap.parse_args()
    from switch_nerf_b200 import synthetic as O
    from switch_nerf_b200.configs import make_hparams
    from switch_nerf_b200.nerf_moe import get_nerf_moe_inner
    from switch_nerf_b200.rendering import render_rays
    dev = torch.device("cuda", 0)
    E, coarse, fine, chunk = 8, 257, 257, 131072
    sd_cpu = O.benchmark_state_dict(num_experts=E, appearance_count=2048, seed=0, n_rays=8192, coarse=coarse)
    sd = {k: v.to(dev) for k, v in sd_cpu.items()}
    rays, idx = O.synthetic_rays(args.rays, 2048, seed=100)
    rays, idx = rays.to(dev), idx.to(dev)
    with torch.no_grad():
        for _ in range(args.warmup):
            rgb, depth = render(sd, rays, idx, coarse, fine, chunk)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            rgb, depth = render(sd, rays, idx, coarse, fine, chunk)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        hp = make_hparams(num_experts=E, capacity_factor=1.0, bpr=True, model_chunk_size=chunk, coarse_samples=coarse,
                          fine_samples=fine, amp_bf16=True, moe_return_gates=False)
        model = get_nerf_moe_inner(hp, 2048, 3)
        model.load_state_dict(sd_cpu)
        model = model.to(dev).eval()
        for _ in range(3):
            ours = render_rays(model, None, rays, idx, hp, None, None, True, True, False)[0]
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            ours = render_rays(model, None, rays, idx, hp, None, None, True, True, False)[0]
        e1.record()
        torch.cuda.synchronize()
        ms_ours = e0.elapsed_time(e1) / args.steps
    samples = args.rays * (coarse + fine)
    d = (ours["rgb_fine"] - rgb).abs()
    print(json.dumps({
        "what": "unfused PyTorch (bf16 autocast, cuBLAS, index_put/gather dispatch) vs switch_nerf_b200, same B200, same rays",
        "workload": f"{args.rays} rays x ({coarse}+{fine}), E={E}, chunk {chunk}, cf 1.0, BPR, balanced gate",
        "torch_unfused": {"ms_per_step": ms, "samples_per_s": samples / (ms * 1e-3)},
        "switch_nerf_b200": {"ms_per_step": ms_ours, "samples_per_s": samples / (ms_ours * 1e-3)},
        "speedup": ms / ms_ours,
        "rgb_max_abs_diff": float(d.max()), "rgb_mean_abs_diff": float(d.mean()),
        "psnr_db": float(-10 * torch.log10((ours["rgb_fine"] - rgb).square().mean())),
        "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
        "note": "no L2 flush between steps in either arm; bench.py is the contract number"}))


if __name__ == "__main__":
    main()
