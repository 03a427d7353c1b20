"""CPU (gloo, world_size 2): the N>1 logic of bench.py -- rays are independent units, ranks take disjoint
seeded shards, no data-path collective; the only collective is the MAX-reduce of the elapsed time."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from switch_nerf_b200 import synthetic as O
    rays, idx = O.synthetic_rays(64, 16, seed=100 + rank)       # bench.py: seed = 100 + rank
    # per-rank elapsed time -> MAX over ranks (what bench.py reports), samples -> SUM over ranks
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = torch.tensor([rays.shape[0] * 514], dtype=torch.int64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    gathered = [torch.zeros_like(rays) for _ in range(world)]
    dist.all_gather(gathered, rays)
    if rank == 0:
        torch.save({"t": t, "n": n, "distinct": not torch.equal(gathered[0], gathered[1])}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    out = str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
    r = torch.load(out)
    assert float(r["t"]) == 2.0            # max over ranks
    assert int(r["n"]) == 2 * 64 * 514     # whole-job samples
    assert r["distinct"]                   # ranks render different ray shards


def test_reference_arm_rank_nonzero_exits_quietly():
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def _ep_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from argparse import Namespace
    from switch_nerf_b200.configs import make_hparams
    from switch_nerf_b200.expert_parallel import gather_handles
    from switch_nerf_b200.nerf_moe import expert_parallel_env, gather_expert_shards, get_nerf_moe_inner
    # reference plumbing (runner.py:100-101, nerf_moe.py:284-289): E/W experts per rank, seeded by rank
    hp = make_hparams(num_experts=4, amp_bf16=True)
    hp.moe_local_expert_num, hp.no_expert_parallel = 2, False
    hp.parallel_env = Namespace(global_rank=rank)
    model = get_nerf_moe_inner(hp, 8, 3)
    moe = model.layers["0"]
    w0 = moe.experts[0].weights[0].detach()
    full = gather_expert_shards(w0)
    mine = [torch.zeros_like(w0) for _ in range(world)]
    dist.all_gather(mine, w0)
    blob = gather_handles(bytes([rank]) * 64)
    res = {"env": expert_parallel_env(hp), "local": moe.num_local_experts, "global": moe.num_global_experts,
           "wg": tuple(moe.gates[0].wg.weight.shape), "w0": tuple(w0.shape), "full": tuple(full.shape),
           "order_ok": torch.equal(full, torch.cat(mine, 0)), "ranks_differ": not torch.equal(mine[0], mine[1]),
           "blob": blob, "desc_E": model._desc().num_experts}
    torch.save(res, out + f".{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_expert_parallel_host_logic(tmp_path):
    """World 2 (gloo): expert-sharded parameters (E/W per rank, rank-seeded like the reference), the gate over the
    global expert count, the all-gather that rebuilds the full expert set for packing, IPC handle exchange order."""
    out = str(tmp_path / "ep")
    mp.spawn(_ep_worker, args=(2, 29537, out), nprocs=2, join=True)
    for rank in range(2):
        r = torch.load(out + f".{rank}", weights_only=False)
        assert r["env"] == (rank, 2)
        assert (r["local"], r["global"], r["desc_E"]) == (2, 4, 4)
        assert r["wg"] == (4, 256) and r["w0"] == (2, 256, 256) and r["full"] == (4, 256, 256)
        assert r["order_ok"] and r["ranks_differ"]
        assert r["blob"] == bytes([0]) * 64 + bytes([1]) * 64
