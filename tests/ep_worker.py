"""Expert-parallel parity worker -- run under torchrun with one rank per GPU (tests/test_gpu_parity.py spawns it).

Every rank renders its own ray shard twice with identical (seeded) weights: all experts local, then with the
experts sharded over the ranks (P2P record exchange, csrc/snb_ep.cu).  SURVEY F5 / §8e: the two must agree
exactly -- capacity and drops are decided per source rank, and launch #2's per-row arithmetic does not depend on
the tile a row lands in.  Prints one JSON line per rank-0 case and exits non-zero on any mismatch.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

# every rank must make the same number of model-chunk calls (rays x samples / chunk, rounded up, per pass)
CASES = [
    # name, rays on rank r, coarse, fine, chunk, experts, capacity factor, bpr, balanced gate
    dict(name="small_e4", rays=lambda r: 300 - 10 * r, coarse=32, fine=32, chunk=4096, E=4, cf=1.0, bpr=True, balance=True),
    dict(name="drops_e8_cf05", rays=lambda r: 256 - 3 * r, coarse=64, fine=0, chunk=4096, E=8, cf=0.5, bpr=True, balance=False),
    dict(name="nobpr_e2", rays=lambda r: 200, coarse=48, fine=16, chunk=2048, E=2, cf=2.0, bpr=False, balance=True),
    dict(name="full_chunk_e8", rays=lambda r: 1024, coarse=257, fine=257, chunk=131072, E=8, cf=1.0, bpr=True, balance=True),
]


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=device)
    from switch_nerf_b200 import synthetic as O
    from switch_nerf_b200.configs import make_hparams
    from switch_nerf_b200.expert_parallel import ExpertParallelGroup
    from switch_nerf_b200.nerf_moe import get_nerf_moe_inner
    from switch_nerf_b200.rendering import render_rays

    failures = 0
    for case in CASES:
        E = case["E"]
        if E % world:
            continue
        n_rays = case["rays"](rank)
        if case["balance"]:
            sd = O.benchmark_state_dict(num_experts=E, appearance_count=64, seed=3, n_rays=256, coarse=case["coarse"])
        else:
            sd = O.synthetic_state_dict(num_experts=E, appearance_count=64, seed=3)
        hp = make_hparams(num_experts=E, capacity_factor=case["cf"], bpr=case["bpr"], model_chunk_size=case["chunk"],
                          coarse_samples=case["coarse"], fine_samples=case["fine"], amp_bf16=True, moe_return_gates=True)
        model = get_nerf_moe_inner(hp, 64, 3)
        model.load_state_dict(sd)
        model = model.to(device).eval()
        rays, idx = O.synthetic_rays(n_rays, 64, seed=500 + rank)
        rays, idx = rays.to(device), idx.to(device)

        def render():
            res = render_rays(model, None, rays, idx, hp, None, None, True, True, False)[0]
            torch.cuda.synchronize()
            return {k: v.clone() for k, v in res.items() if torch.is_tensor(v)}

        x = torch.cat([torch.rand(1000 + rank, 3, device=device) - 0.5,
                       torch.nn.functional.normalize(torch.randn(1000 + rank, 3, device=device), dim=1),
                       torch.randint(0, 64, (1000 + rank, 1), device=device).float()], 1)
        noise = torch.randn(1000 + rank, 1, device=device)

        local = render()
        local_fwd = model(x, sigma_noise=noise)["outputs"].clone()
        group = ExpertParallelGroup(E, case["chunk"], case["cf"])
        group.attach(model)
        ep = render()
        ep2 = render()                    # buffer sets and epochs are reused across calls
        ep_fwd = model(x, sigma_noise=noise)["outputs"].clone()
        torch.cuda.synchronize()
        dist.barrier()
        group.detach(model)
        again = render()
        report = {"case": case["name"], "rank": rank, "world": world, "keys": sorted(local)}
        bad = {}
        for k in local:
            for tag, other in (("ep", ep), ("ep_repeat", ep2), ("detached", again)):
                if not torch.equal(local[k], other[k]):
                    a, b = local[k].double(), other[k].double()
                    bad[f"{tag}:{k}"] = {"max_abs": float((a - b).abs().max()),
                                         "mismatched": int((a != b).sum()), "numel": int(a.numel())}
        if not torch.equal(local_fwd, ep_fwd):
            d = (local_fwd - ep_fwd).abs()
            bad["forward_noise"] = {"max_abs": float(d.max()), "mismatched_rows": int((d.amax(1) > 0).sum()),
                                    "rows": int(d.shape[0])}
        report["mismatches"] = bad
        flag = torch.tensor([len(bad)], device=device)
        dist.all_reduce(flag)
        failures += int(flag.item())
        if bad or rank == 0:
            print(json.dumps(report), flush=True)
        dist.barrier()
        group.close()
        model.release()
    dist.barrier()
    dist.destroy_process_group()
    if failures:
        raise SystemExit(1)
    if rank == 0:
        print("EP_PARITY_OK", flush=True)


if __name__ == "__main__":
    main()
