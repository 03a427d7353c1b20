"""TEST INFRASTRUCTURE ONLY -- goldens of the UNMODIFIED reference under `torch.autocast("cuda", bfloat16)`.

Run ON THE GPU BOX (the reference copy travels in the git-ignored baseline/_ref, oracle/install_ref.py):

    python -m oracle.make_golden_cuda --out gpurun_out/golden_cuda [--device cuda]

The benchmarked precision of switch_nerf_b200 is the reference's README training recipe (`--amp_use_bfloat16`):
nn.Linear / baddbmm in bf16 with fp32 accumulation, LayerNorm / gate / softmax / softplus in fp32.  That rounding
map only exists on a CUDA device (CPU autocast keeps LayerNorm and softplus in bf16), so the fixtures that pin it
have to be produced there: the reference's own `NeRFMoE` / `MipNeRFMoE` / `rendering.render_rays` /
`rendering_mip.render_rays`, unmodified, with the Tutel shims' torch ops running on the GPU (oracle/ref_shims.py),
`argsort(stable=True)` and TF32 off.  Weights and inputs are regenerated from seeds
(switch_nerf_b200.synthetic); only outputs are stored.  The files are then committed under tests/golden/ and
checked by tests/test_oracle_golden.py (oracle, flavor="cuda", on CPU) and tests/test_gpu_parity.py (the
tcgen05 path on the GPU).

`--device cpu` runs the same code under CPU autocast (a dry run of the script where there is no GPU).
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shims as R            # noqa: E402
from oracle import switch_nerf_oracle as O   # noqa: E402
from oracle.make_golden import model_inputs, sd_checksum   # noqa: E402
from switch_nerf_b200 import synthetic as SY  # noqa: E402

warnings.filterwarnings("ignore")


def bf16_ulp(x: torch.Tensor) -> torch.Tensor:
    """Spacing of bf16 numbers at |x| (8 significant bits)."""
    e = torch.floor(torch.log2(x.abs().clamp_min(2.0 ** -126)))
    return torch.pow(2.0, e - 7)


def ulp_histogram(a: torch.Tensor, b: torch.Tensor) -> dict:
    """Histogram of |a-b| in units of the bf16 ulp of b (for bf16-valued outputs such as sigmoid(rgb))."""
    d = ((a - b).abs() / bf16_ulp(b)).round().long().view(-1)
    n = d.numel()
    return {"0": float((d == 0).sum()) / n, "1": float((d == 1).sum()) / n, "2": float((d == 2).sum()) / n,
            ">2": float((d > 2).sum()) / n}


class GateTap:
    """Records the fp32 gates the reference hands to extract_critical (tutel_moe_layer_nobatch.py:134) and what
    it returns, without touching the reference: the module-level name is wrapped for the duration of a call."""

    def __init__(self, no_batch=False):
        import switch_nerf.modules.tutel_moe_ext.tutel_moe_layer_nobatch as M
        self.M, self.name = M, ("extract_critical_nobatch" if no_batch else "extract_critical")
        self.orig = getattr(M, self.name)
        self.calls = []

    def __enter__(self):
        def wrapped(gates, *a, **k):
            r = self.orig(gates, *a, **k)
            crit = r[0]
            self.calls.append({"gates": gates.detach().float().cpu(), "idx": crit[1][0].detach().cpu(),
                               "loc": crit[2][0].detach().cpu(),
                               "cap": (1 << 30) if self.name.endswith("nobatch") else int(crit[-1])})
            return r
        setattr(self.M, self.name, wrapped)
        return self

    def __exit__(self, *exc):
        setattr(self.M, self.name, self.orig)


def run_model(m, x, device, no_batch=False):
    with torch.no_grad(), R.stable_argsort(), GateTap(no_batch) as tap, torch.autocast(device.type, dtype=torch.bfloat16):
        r = m(x.to(device))
    out = r["outputs"].float().cpu()
    idx = r["extras"]["moe_gates"][0].view(-1).to(torch.int32).cpu()
    l_aux = r["extras"]["moe_loss"].float().cpu()
    return out, idx, l_aux, tap.calls[0]


def oracle_report(tag, x, sd, cfg, out, idx, tap, flavor):
    """Restatement (CPU, mode bf16, given flavor) vs the reference run: routing agreement, ulp histogram of rgb on
    same-route samples, sigma error.  This is what pins oracle flavor="cuda"."""
    o2, ex = O.nerf_moe_forward(x, sd, cfg, mode="bf16", flavor=flavor)
    ocap = (1 << 30) if cfg.get("moe_no_batch") else ex["capacity"]
    same = (ex["idx"].to(torch.int32) == idx) & ((ex["loc"] < ocap) == (tap["loc"] < tap["cap"]))
    rep = {"tag": tag, "samples": int(x.shape[0]), "route_agree": float(same.float().mean()),
           "rgb_ulp_hist_same_route": ulp_histogram(o2[same, :3], out[same, :3]),
           "rgb_max_abs_same_route": float((o2[same, :3] - out[same, :3]).abs().max()),
           "sigma_max_abs_same_route": float((o2[same, 3] - out[same, 3]).abs().max()),
           "sigma_max_rel_same_route": float(((o2[same, 3] - out[same, 3]).abs() / out[same, 3].abs().clamp_min(1e-3)).max()),
           "mean_abs_all": float((o2 - out).abs().mean()),
           "gates_max_abs": float((ex["gates"] - tap["gates"]).abs().max())}
    print(json.dumps(rep))
    return rep


def save_model_case(outdir, tag, params, sd, x, out, idx, l_aux, tap, store_x=True):
    g = tap["gates"]
    top2 = torch.topk(g, min(2, g.shape[1]), dim=1).values
    margin = (top2[:, 0] - top2[:, -1]) if g.shape[1] > 1 else top2[:, 0]
    d = dict(params=np.array(params, dtype=np.float64), sd_checksum=np.array([sd_checksum(sd)]),
             outputs=out.numpy(), idx=idx.numpy().astype(np.int8), loc=tap["loc"].numpy().astype(np.int32),
             l_aux=l_aux.numpy().reshape(-1), capacity=np.array([tap["cap"]]),
             gate_top=top2[:, 0].numpy(), gate_margin=margin.numpy().astype(np.float16))
    if store_x:
        d["x"] = x.numpy()
    np.savez_compressed(os.path.join(outdir, f"model_{tag}.npz"), **d)


def golden_model(outdir, device, tag, E, cf, bpr, S, flavor, no_batch=False, gate_scale=4.0, seed=3, width=256, mip=False):
    count = 16
    sd = SY.synthetic_state_dict(num_experts=E, appearance_count=count, seed=seed, gate_scale=gate_scale, width=width)
    hp = R.make_hparams(num_experts=E, capacity_factor=cf, bpr=bpr, model_chunk_size=4096, coarse_samples=32,
                        fine_samples=32, amp_bf16=True, width=width,
                        moe_expert_type="seqexperts" if no_batch else "expertmlp",
                        nerfmoe_class_name="MipNeRFMoE" if mip else "NeRFMoE")
    m = R.build_reference_model(hp, appearance_count=count, xyz_dim=3).eval()
    if no_batch:
        from switch_nerf.models.model_utils import convert_to_seqexperts
        m.load_state_dict({k.replace("module.", ""): v for k, v in convert_to_seqexperts({k: v.clone() for k, v in sd.items()}).items()})
        m.set_no_batch(True)
    else:
        m.load_state_dict(sd)
    m = m.to(device)
    x = model_inputs(S, count, seed + 100)
    if mip:   # [mean3, cov_diag3, dir3, idx]: covariances of the Mission-Bay scale (radii 5e-4..2e-3, t in 0.01..10)
        g = torch.Generator().manual_seed(seed + 200)
        cov = torch.rand(S, 3, generator=g) ** 4 * 1e-2
        x = torch.cat([x[:, :3], cov, x[:, 3:]], 1)
    out, idx, l_aux, tap = run_model(m, x, device, no_batch)
    cfg = O.default_cfg(sd, cf, bpr, moe_no_batch=no_batch, mip=mip)
    rep = oracle_report(tag, x, sd, cfg, out, idx, tap, flavor)
    save_model_case(outdir, tag, [E, cf, int(bpr), S, seed, gate_scale, count, int(no_batch), 1, width, int(mip)], sd, x, out,
                    idx, l_aux, tap)
    return rep


def bench_inputs(n_rays=8192, coarse=257):
    sd = SY.benchmark_state_dict(num_experts=8, appearance_count=2048, seed=0, n_rays=n_rays, coarse=coarse)
    rays, idx = SY.synthetic_rays(n_rays, 2048, seed=100)
    return sd, rays, idx


def bench_chunk_x(rays, idx, coarse, rows):
    """First `rows` rows of the coarse pass of the benchmark's ray batch, built as rendering.py:85-90, 306-314, 357-362."""
    n = (rows + coarse - 1) // coarse
    t = torch.linspace(0, 1, coarse)
    z = rays[:n, 6:7] * (1 - t) + rays[:n, 7:8] * t
    xyz = rays[:n, None, 0:3] + rays[:n, None, 3:6] * z[..., None]
    x = torch.cat([xyz.reshape(-1, 3), rays[:n, None, 3:6].expand(n, coarse, 3).reshape(-1, 3),
                   idx[:n, None, None].expand(n, coarse, 1).reshape(-1, 1).float()], 1)
    return x[:rows].contiguous()


def golden_bench_chunk(outdir, device, flavor, rows=131072):
    """One full Building model chunk (S = 131072, E = 8, cf = 1, BPR) with the benchmark's weights and rays."""
    sd, rays, idx = bench_inputs()
    hp = R.make_hparams(num_experts=8, capacity_factor=1.0, bpr=True, model_chunk_size=131072, amp_bf16=True)
    m = R.build_reference_model(hp, appearance_count=2048).eval()
    m.load_state_dict(sd)
    m = m.to(device)
    x = bench_chunk_x(rays, idx, 257, rows)
    out, gi, l_aux, tap = run_model(m, x, device)
    rep = oracle_report("bench_chunk", x, sd, O.default_cfg(sd, 1.0, True), out, gi, tap, flavor)
    save_model_case(outdir, "bench_chunk_bf16cuda", [8, 1.0, 1, rows, 0, 4.0, 2048, 0, 1, 256, 0], sd, x, out, gi, l_aux, tap,
                    store_x=False)
    return rep


def golden_render(outdir, device, tag, sd, rays, idx, hp, time_it=False):
    from switch_nerf import rendering
    count = sd["embedding_a.weight"].shape[0]
    m = R.build_reference_model(hp, appearance_count=count).eval()
    m.load_state_dict(sd)
    m = m.to(device)
    rays_d, idx_d = rays.to(device), idx.to(device)

    def step():
        with torch.no_grad(), R.stable_argsort(), torch.autocast(device.type, dtype=torch.bfloat16):
            return rendering.render_rays(m, None, rays_d, idx_d, hp, None, None, True, True, False)[0]

    res = step()
    timing = None
    if time_it and device.type == "cuda":
        for _ in range(2):
            step()
        ts = []
        for _ in range(5):
            torch.cuda.synchronize(); t0 = time.perf_counter(); step(); torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        timing = {"ms_per_step_median": 1e3 * ts[len(ts) // 2],
                  "samples_per_s": rays.shape[0] * (hp.coarse_samples + hp.fine_samples) / ts[len(ts) // 2]}
        print(json.dumps({"tag": tag, "reference_on_gpu": timing}))
    save = {k: v.float().cpu().numpy() for k, v in res.items() if not k.startswith("moe_gates")}
    for k in ("moe_gates_coarse", "moe_gates_fine"):
        if k in res:
            save[k] = res[k].cpu().numpy().astype(np.int8)
    np.savez_compressed(os.path.join(outdir, f"render_{tag}.npz"),
                        params=np.array([sd["layers.0.gates.0.wg.weight"].shape[0], hp.moe_capacity_factor,
                                         int(hp.batch_prioritized_routing), rays.shape[0], hp.coarse_samples, hp.fine_samples,
                                         hp.model_chunk_size], dtype=np.float64),
                        sd_checksum=np.array([sd_checksum(sd)]), **save)
    return res, timing


def golden_render_mip(outdir, device, tag, E, width, n_rays, cs, fs, chunk, gate_scale=3.0, seed=9):
    from switch_nerf import rendering_mip
    count = 16
    sd = SY.synthetic_state_dict(num_experts=E, appearance_count=count, seed=seed, gate_scale=gate_scale, width=width)
    hp = R.make_hparams(num_experts=E, model_chunk_size=chunk, coarse_samples=cs, fine_samples=fs, width=width,
                        nerfmoe_class_name="MipNeRFMoE", amp_bf16=True)
    hp.perturb = 0
    m = R.build_reference_model(hp, appearance_count=count).eval()
    m.load_state_dict(sd)
    m = m.to(device)
    rays, idx = SY.synthetic_rays(n_rays, count, seed=seed + 1)
    g = torch.Generator().manual_seed(seed + 2)
    radii = torch.rand(n_rays, 1, generator=g) * 1.5e-3 + 5e-4
    with torch.no_grad(), R.stable_argsort(), torch.autocast(device.type, dtype=torch.bfloat16):
        res, _ = rendering_mip.render_rays(m, rays.to(device), radii.to(device), idx.to(device), hp, True, True)
    save = {k: v.float().cpu().numpy() for k, v in res.items() if not k.startswith("moe_gates")}
    for k in ("moe_gates_coarse", "moe_gates_fine"):
        save[k] = res[k].cpu().numpy().astype(np.int8)
    np.savez_compressed(os.path.join(outdir, f"render_{tag}.npz"),
                        params=np.array([E, width, n_rays, cs, fs, chunk, seed, gate_scale, count], dtype=np.float64),
                        sd_checksum=np.array([sd_checksum(sd)]), rays=rays.numpy(), radii=radii.numpy(),
                        image_indices=idx.numpy(), **save)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/golden_cuda")
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--quick", action="store_true", help="small sizes (dry run of the script)")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    device = torch.device(a.device)
    flavor = "cuda" if device.type == "cuda" else "cpu"
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    R.install_shims()
    reports = []
    # per-sample model goldens (same seeds / inputs as the CPU-autocast fixture model_e8_cf1_bpr_bf16cpu)
    reports.append(golden_model(a.out, device, "e8_cf1_bpr_bf16cuda", 8, 1.0, True, 4096, flavor))
    reports.append(golden_model(a.out, device, "e8_cf05_nobpr_bf16cuda", 8, 0.5, False, 5000, flavor))
    reports.append(golden_model(a.out, device, "e8_cf2_bpr_bf16cuda", 8, 2.0, True, 4096, flavor, seed=4))
    reports.append(golden_model(a.out, device, "e4_cf1_bpr_bf16cuda", 4, 1.0, True, 4096, flavor))
    reports.append(golden_model(a.out, device, "e4_nobatch_bf16cuda", 4, 1.0, False, 3000, flavor, no_batch=True))
    reports.append(golden_model(a.out, device, "mip_e8_w512_bf16cuda", 8, 1.0, True, 1024 if a.quick else 4096, flavor,
                                width=512, mip=True, gate_scale=3.0, seed=9))
    reports.append(golden_model(a.out, device, "mip_e4_w256_bf16cuda", 4, 1.0, True, 1024 if a.quick else 4096, flavor,
                                width=256, mip=True, gate_scale=3.0, seed=9))
    # one full Building chunk with the benchmark's weights (131072 rows)
    reports.append(golden_bench_chunk(a.out, device, flavor, rows=8192 if a.quick else 131072))
    # renders: BASELINE.json configs[0] and the benchmark configuration itself (configs[1])
    sd1 = SY.synthetic_state_dict(num_experts=4, appearance_count=16, seed=5, gate_scale=4.0)
    rays1, idx1 = SY.synthetic_rays(256, 16, seed=6)
    hp1 = R.make_hparams(num_experts=4, capacity_factor=1.0, bpr=True, model_chunk_size=4096, coarse_samples=32,
                         fine_samples=32, amp_bf16=True)
    golden_render(a.out, device, "config1_bf16cuda", sd1, rays1, idx1, hp1)
    sdb, raysb, idxb = bench_inputs()
    nb = 64 if a.quick else 8192
    hpb = R.make_hparams(num_experts=8, capacity_factor=1.0, bpr=True, model_chunk_size=131072, coarse_samples=257,
                         fine_samples=257, amp_bf16=True, moe_return_gates=False)
    _, timing = golden_render(a.out, device, "bench_building_bf16cuda", sdb, raysb[:nb], idxb[:nb], hpb, time_it=True)
    golden_render_mip(a.out, device, "mip_w256_bf16cuda", 4, 256, 128, 33, 33, 3000)
    golden_render_mip(a.out, device, "mip_mission_bay_w512_bf16cuda", 8, 512, 64, 33, 33, 1500)
    json.dump({"device": str(device), "torch": torch.__version__,
               "gpu": torch.cuda.get_device_name(0) if device.type == "cuda" else None,
               "oracle_vs_reference": reports, "reference_on_gpu_bench_building": timing},
              open(os.path.join(a.out, "report.json"), "w"), indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
