"""TEST INFRASTRUCTURE (oracle tooling for the NEXT scope row, SURVEY §8f-1: backward of the fused path).

Runs one training-style step of the UNMODIFIED reference on CPU (through oracle/ref_shims.py) and of the oracle
restatement, with the reference's loss (runner.py:1100-1111, 646-651):

    loss = mse(rgb_fine, target) + moe_l_aux_wt * (mean(gate_loss_fine) + mean(gate_loss_coarse)) / 2

and writes a compact digest of the reference's parameter gradients to tests/golden/grad_config1.npz.  Loss and
gradients of the two are bit-identical on CPU -- provided the restatement cuts the same paths (the fine samples are
drawn from DETACHED coarse weights, rendering.py:240; this fixture caught the oracle missing that detach).  A CUDA
backward will not be bit-identical (different summation orders; torch.cumprod's backward divides by
1 - alpha + 1e-8 = 1e-8 at the last sample): grad_close() is the tolerance meant for it.

    python -m oracle.make_golden_grad
"""
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import ref_shims as R
from oracle import switch_nerf_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASE = dict(E=4, cf=1.0, bpr=True, n_rays=64, cs=16, fs=16, chunk=512, seed=3, gate_scale=3.0, count=16, wt=5e-4)


def training_loss(res, target, wt):
    gate_loss = (res["gate_loss_fine"].mean() + res["gate_loss_coarse"].mean()) / 2.0      # runner.py:1105-1111
    return F.mse_loss(res["rgb_fine"], target) + wt * gate_loss                           # :1100-1102, 646-651


def case_inputs(c=CASE):
    sd = O.synthetic_state_dict(num_experts=c["E"], appearance_count=c["count"], seed=c["seed"], gate_scale=c["gate_scale"])
    rays, idx = O.synthetic_rays(c["n_rays"], c["count"], seed=c["seed"] + 1)
    g = torch.Generator().manual_seed(c["seed"] + 6)
    target = torch.rand(c["n_rays"], 3, generator=g)
    return sd, rays, idx, target


def reference_grads(c=CASE):
    R.install_shims()
    from switch_nerf import rendering
    sd, rays, idx, target = case_inputs(c)
    hp = R.make_hparams(num_experts=c["E"], capacity_factor=c["cf"], bpr=c["bpr"], model_chunk_size=c["chunk"],
                        coarse_samples=c["cs"], fine_samples=c["fs"])
    m = R.build_reference_model(hp, appearance_count=c["count"]).eval()      # eval: deterministic sampling, same math
    m.load_state_dict(sd)
    with R.stable_argsort():
        res, _ = rendering.render_rays(m, None, rays, idx, hp, None, None, True, True, False)
        loss = training_loss(res, target, c["wt"])
        loss.backward()
    return float(loss.detach()), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}


def oracle_grads(c=CASE):
    sd, rays, idx, target = case_inputs(c)
    sdg = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    res = O.render_rays(sdg, O.default_cfg(sd, c["cf"], c["bpr"]), rays, idx, coarse_samples=c["cs"], fine_samples=c["fs"],
                        model_chunk_size=c["chunk"])
    loss = training_loss(res, target, c["wt"])
    loss.backward()
    return float(loss.detach()), {k: v.grad.clone() for k, v in sdg.items() if v.grad is not None}


def sample_of(t, n=512):
    flat = t.reshape(-1)
    stride = max(1, flat.numel() // n)
    return flat[::stride][:n].clone()


def grad_close(a, b, scale):
    """|a - b| <= 1e-4 * (largest gradient magnitude of the whole model) + 5 % of this tensor's own largest entry."""
    return float((a - b).abs().max()) <= 1e-4 * scale + 5e-2 * float(b.abs().max())


def main():
    torch.set_num_threads(8)
    loss_ref, g_ref = reference_grads()
    loss_ora, g_ora = oracle_grads()
    assert loss_ref == loss_ora, (loss_ref, loss_ora)
    scale = max(float(v.abs().max()) for v in g_ref.values())
    for k, v in g_ref.items():
        assert grad_close(g_ora[k], v, scale), k
    # small fixture: per tensor [sum, l1, linf] + a strided sample of <= 512 entries (the full set is 7.6 MB)
    save = {}
    worst = 0.0
    for k, v in g_ref.items():
        flat = v.reshape(-1).double()
        save["stats/" + k] = np.array([float(flat.sum()), float(flat.abs().sum()), float(flat.abs().max())])
        save["sample/" + k] = sample_of(v).numpy()
        worst = max(worst, float((g_ora[k] - v).abs().max()) / scale)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "grad_config1.npz"), loss=np.array([loss_ref]), scale=np.array([scale]),
                        case=np.array([CASE[k] for k in ("E", "cf", "bpr", "n_rays", "cs", "fs", "chunk", "seed", "gate_scale", "count", "wt")],
                                      dtype=np.float64), **save)
    print("loss", loss_ref, "grad scale", scale, len(g_ref), "tensors; worst |oracle - reference| / scale =", worst)


if __name__ == "__main__":
    main()
