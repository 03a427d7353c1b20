"""TEST INFRASTRUCTURE (oracle for the NEXT scope row): the backward of the hot path written out stage
by stage, the way CUDA kernels will compute it -- explicit formulas, no autograd -- and checked against autograd of the
pinned forward restatement (tests/test_oracle_golden.py::test_backward_plan_*).

What the reference differentiates (fp32, batched dispatch, bg_nerf=None):
  rendering.py:436-494   composite: rgb = sum_i w_i c_i, w_i = alpha_i T_i, T exclusive cumprod of (1 - alpha + 1e-8);
                         depth / depth_variance are computed under no_grad (:479-494)
  rendering.py:419-431   merge of fine and coarse samples by sort + gather (gradient = scatter back through `ordering`)
  rendering.py:240       fine samples drawn from DETACHED coarse weights (no gradient through the sample positions)
  nerf_moe.py:320-455    heads, layer "1"/"2", sigma (softplus(x-1)), embedding_a lookup, xyz layer, external gate MLP,
                         LayerNorm
  tutel_fast_dispatch.py:15-78   GatingEncoder / GatingDecoder backward: dispatch^T, combine^T and the gate-value
                         gradient (K6: d gate[s] = <dy[s], expert_out[row(s)]>)
  tutel_fast_dispatch.py:141-145 l_aux = E/S^2 * sum_e me_e ce_e, me_e = sum_s gates[s,e] (ce is a count: no gradient)
Positional encodings have no parameters and the sample positions are inputs, so nothing flows into x.

Each function returns plain tensors; `model_chunk_backward` needs only the chunk input, the weights and the upstream
gradients (it recomputes the forward intermediates, as a fused kernel would)."""
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from oracle import switch_nerf_oracle as O

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------------------------
# ray side
# --------------------------------------------------------------------------------------------------------------
def composite_backward(z_vals: Tensor, rgbs: Tensor, sigmas: Tensor, last_delta: Tensor, d_rgb: Tensor
                       ) -> Tuple[Tensor, Tensor]:
    """d(loss)/d(rgbs) [N,S,3] and d(loss)/d(sigmas) [N,S] given d(loss)/d(rgb_ray) [N,3].

    With q_j = 1 - alpha_j + 1e-8, T_i = prod_{j<i} q_j, w_i = alpha_i T_i:
        dL/dc_i     = w_i * dL/drgb
        g_i        := dL/dw_i = <c_i, dL/drgb>
        dL/dalpha_j = g_j T_j - (1/q_j) * sum_{i>j} g_i w_i            (a reverse exclusive scan of g_i w_i)
        dL/dsigma_j = dL/dalpha_j * delta_j * exp(-delta_j sigma_j)
    The 1/q_j factor is what torch.cumprod's autograd also divides by; q_j = 1e-8 only at the last sample, where the
    suffix sum is empty, so the kernel never divides by it."""
    deltas = torch.cat([z_vals[:, 1:] - z_vals[:, :-1], last_delta], -1)
    e = torch.exp(-deltas * sigmas)
    alphas = 1 - e
    q = 1 - alphas + 1e-8
    T = torch.cumprod(q, -1)
    T = torch.cat((torch.ones_like(T[..., :1]), T[..., :-1]), -1)
    w = alphas * T
    d_rgbs = w.unsqueeze(-1) * d_rgb.unsqueeze(1)
    g = (rgbs * d_rgb.unsqueeze(1)).sum(-1)
    gw = g * w
    suffix = torch.flip(torch.cumsum(torch.flip(gw, [-1]), -1), [-1]) - gw          # sum_{i>j} g_i w_i
    d_alpha = g * T - suffix / q
    d_sigmas = d_alpha * deltas * e
    return d_rgbs, d_sigmas


def merge_backward(order: Tensor, n_fine: int, d_rgbs: Tensor, d_sigmas: Tensor):
    """rendering.py:419-431 backward: the merged sample k came from position order[k] of cat([fine, coarse])."""
    n = order.shape[0]
    total = order.shape[1]
    dr = torch.zeros(n, total, 3, dtype=d_rgbs.dtype).scatter_(1, order.unsqueeze(-1).expand(-1, -1, 3), d_rgbs)
    ds = torch.zeros(n, total, dtype=d_sigmas.dtype).scatter_(1, order, d_sigmas)
    return (dr[:, :n_fine], ds[:, :n_fine]), (dr[:, n_fine:], ds[:, n_fine:])


# --------------------------------------------------------------------------------------------------------------
# model chunk
# --------------------------------------------------------------------------------------------------------------
def _lin_bwd(x: Tensor, w: Tensor, dy: Tensor):
    """y = x W^T + b  ->  dx = dy W, dW = dy^T x, db = sum dy."""
    return dy @ w, dy.t() @ x, dy.sum(0)


def model_chunk_backward(x: Tensor, sd: Dict[str, Tensor], cfg: dict, d_out: Tensor, d_l_aux: float
                         ) -> Dict[str, Tensor]:
    """Parameter gradients of one model chunk (nerf_moe.py:320-455 + MoE layer), fp32, batched dispatch.
    d_out [S,4] = d(loss)/d([rgb, sigma]); d_l_aux = d(loss)/d(l_aux of this chunk)."""
    assert not cfg.get("mip") and not cfg.get("moe_no_batch", False)
    E, L, skips = cfg["num_experts"], cfg["expert_layers"], cfg["skips"] or []
    S = x.shape[0]
    grads: Dict[str, Tensor] = {}
    # ---------------- forward recomputation (what launch #1 / #2 keep on chip) ----------------
    pe = O.embedding(x[:, :3], cfg["pos_xyz_dim"])
    Wx, bx = sd["layers.xyz.fcs.0.weight"], sd["layers.xyz.fcs.0.bias"]
    h = F.linear(pe, Wx, bx)
    gate_acts = [h]                                   # inputs of the gate MLP layers
    t = h
    ng = cfg["gate_layers"]
    for i in range(ng):
        t = F.linear(t, sd[f"layers.moe_external_gate.fcs.{i}.weight"], sd[f"layers.moe_external_gate.fcs.{i}.bias"])
        if i < ng - 1:
            t = F.relu(t)
            gate_acts.append(t)
    g = t
    lw, lb = sd["layers.gate_input_norm.weight"], sd["layers.gate_input_norm.bias"]
    mean = g.mean(1, keepdim=True)
    var = g.var(1, unbiased=False, keepdim=True)
    rstd = torch.rsqrt(var + 1e-5)
    ghat = (g - mean) * rstd
    gi = ghat * lw + lb
    wg = sd["layers.0.gates.0.wg.weight"]
    gates = F.softmax(F.linear(gi, wg), 1)
    idx, loc, gate_val, cap, _ = O.route_top1(gates, cfg["capacity_factor"], cfg["bpr"])
    idx_l, keep = idx.long(), loc < cap
    rows = idx_l * cap + loc.long()
    buf = torch.zeros(E * cap, h.shape[1])
    buf[rows[keep]] = h[keep]
    Ws = [sd[f"layers.0.experts.0.weights.{j}"] for j in range(L)]          # [E, in, out]
    Bs = [sd[f"layers.0.experts.0.bias.{j}"] for j in range(L)]             # [E, 1, out]
    xin = buf.view(E, cap, -1)
    acts = [xin]                                     # input of expert layer j
    pre = []                                         # pre-activation (after the skip add) of layer j
    skip_src = xin
    t = xin
    for j in range(L):
        t = torch.baddbmm(Bs[j], t, Ws[j])
        if j in skips:
            t = t + skip_src
        pre.append(t)
        if j < L - 1:
            t = F.relu(t)
            if j in skips:
                skip_src = t
            acts.append(t)
    out_rows = t.reshape(E * cap, -1)
    y = torch.where(keep.unsqueeze(1), out_rows[rows.clamp(0, E * cap - 1)] * gate_val.unsqueeze(1), torch.zeros(S, h.shape[1]))
    hr = F.relu(y)
    Wsig, W1 = sd["layers.sigma.fcs.0.weight"], sd["layers.1.fcs.0.weight"]
    sig_pre = F.linear(hr, Wsig, sd["layers.sigma.fcs.0.bias"])
    h1 = F.linear(hr, W1, sd["layers.1.fcs.0.bias"])
    d_pe = O.embedding(x[:, 3:6], cfg["pos_dir_dim"])
    ai = x[:, -1].long()
    cat = torch.cat([h1, d_pe, sd["embedding_a.weight"][ai]], -1)
    W2, Wc = sd["layers.2.fcs.0.weight"], sd["layers.color.fcs.0.weight"]
    h2 = F.relu(F.linear(cat, W2, sd["layers.2.fcs.0.bias"]))
    rgb = torch.sigmoid(F.linear(h2, Wc, sd["layers.color.fcs.0.bias"]))
    # ---------------- backward ----------------
    d_rgb, d_sigma = d_out[:, :3], d_out[:, 3:4]
    d_cpre = d_rgb * rgb * (1 - rgb)                                                     # sigmoid'
    d_h2, grads["layers.color.fcs.0.weight"], grads["layers.color.fcs.0.bias"] = _lin_bwd(h2, Wc, d_cpre)
    d_h2pre = d_h2 * (h2 > 0)
    d_cat, grads["layers.2.fcs.0.weight"], grads["layers.2.fcs.0.bias"] = _lin_bwd(cat, W2, d_h2pre)
    M = h.shape[1]
    d_h1 = d_cat[:, :M]
    d_emb = d_cat[:, M + d_pe.shape[1]:]
    grads["embedding_a.weight"] = torch.zeros_like(sd["embedding_a.weight"]).index_add_(0, ai, d_emb)
    # softplus(x - 1, beta=1, threshold=20): derivative sigmoid(x - 1), 1 above the threshold
    z = sig_pre - 1
    d_sigpre = d_sigma * torch.where(z > 20, torch.ones_like(z), torch.sigmoid(z))
    d_hr_s, grads["layers.sigma.fcs.0.weight"], grads["layers.sigma.fcs.0.bias"] = _lin_bwd(hr, Wsig, d_sigpre)
    d_hr_1, grads["layers.1.fcs.0.weight"], grads["layers.1.fcs.0.bias"] = _lin_bwd(hr, W1, d_h1)
    d_y = (d_hr_s + d_hr_1) * (y > 0)
    # combine^T (GatingDecoder.backward): rows of kept samples get gate * dy; gate value gets <dy, expert_out>
    d_outrows = torch.zeros(E * cap, M)
    d_outrows[rows[keep]] = (d_y * gate_val.unsqueeze(1))[keep]
    d_gate_val = torch.where(keep, (d_y * out_rows[rows.clamp(0, E * cap - 1)]).sum(1), torch.zeros(S))
    # expert stack, layer by layer from the top; the skip connection adds the gradient of layer `skip` to its source
    d_t = d_outrows.view(E, cap, M)
    d_skip = None
    for j in reversed(range(L)):
        if j < L - 1:
            d_t = d_t * (pre[j] > 0)
        if j in skips:
            d_skip = d_t                               # d(pre_j)/d(skip_src) = I
        grads[f"layers.0.experts.0.weights.{j}"] = torch.bmm(acts[j].transpose(1, 2), d_t)
        grads[f"layers.0.experts.0.bias.{j}"] = d_t.sum(1, keepdim=True)
        d_t = torch.bmm(d_t, Ws[j].transpose(1, 2))
    if d_skip is not None:
        assert skips == [s for s in skips if True] and len(skips) == 1, "single skip (building.yaml / mission_bay.yaml)"
        d_t = d_t + d_skip                             # the skip source is the stack input (no earlier skip)
    # dispatch^T (GatingEncoder.backward)
    d_h = torch.zeros(S, M)
    d_h[keep] = d_t.reshape(E * cap, M)[rows[keep]]
    # gate: selected-gate gradient + load-balance term  d l_aux / d gates[s,e] = E/S^2 * ce_e
    ce = torch.bincount(idx_l, minlength=E).to(gates.dtype)
    d_gates = torch.zeros_like(gates)
    d_gates[torch.arange(S), idx_l] = d_gate_val
    d_gates = d_gates + d_l_aux * (E / (S * S)) * ce.unsqueeze(0)
    d_logits = gates * (d_gates - (d_gates * gates).sum(1, keepdim=True))               # softmax'
    d_gi, grads["layers.0.gates.0.wg.weight"], _ = _lin_bwd(gi, wg, d_logits)
    grads["layers.gate_input_norm.weight"] = (d_gi * ghat).sum(0)
    grads["layers.gate_input_norm.bias"] = d_gi.sum(0)
    d_ghat = d_gi * lw
    d_g = rstd * (d_ghat - d_ghat.mean(1, keepdim=True) - ghat * (d_ghat * ghat).mean(1, keepdim=True))   # LayerNorm'
    d_t = d_g
    for i in reversed(range(ng)):
        if i < ng - 1:
            d_t = d_t * (gate_acts[i + 1] > 0)
        d_t, grads[f"layers.moe_external_gate.fcs.{i}.weight"], grads[f"layers.moe_external_gate.fcs.{i}.bias"] = \
            _lin_bwd(gate_acts[i], sd[f"layers.moe_external_gate.fcs.{i}.weight"], d_t)
    d_h = d_h + d_t
    _, grads["layers.xyz.fcs.0.weight"], grads["layers.xyz.fcs.0.bias"] = _lin_bwd(pe, Wx, d_h)
    return grads
