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
