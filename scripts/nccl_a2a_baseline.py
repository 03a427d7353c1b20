"""NCCL comparator for the expert-parallel exchange (north_star: "NCCL only as the baseline comparator").

Times what the reference's EP path would put on the wire for the bench workload: per model chunk two
all_to_all_single calls of the dispatched activations [W, E_local, cap, 256] bf16
(tutel_moe_layer_nobatch.py:164-218), 34 chunks per 8192 x (257+257) step.  Only the
collectives are timed (no permute, no experts) -- the floor of an NCCL-based exchange, to set beside
`bench.py --parallelism ep` minus `--parallelism dp`.   torchrun --nproc-per-node N scripts/nccl_a2a_baseline.py"""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    E, M, chunk, rays, samples = 8, 256, 131072, 8192, 514
    cap = chunk // E
    n_chunks = 2 * -(-(rays * (samples // 2)) // chunk)
    send = torch.randn(world, E // world, cap, M, device=dev).to(torch.bfloat16)
    recv = torch.empty_like(send)
    for _ in range(5):
        dist.all_to_all_single(recv, send)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 5
    e0.record()
    for _ in range(steps * n_chunks):
        dist.all_to_all_single(recv, send)      # dispatch
        dist.all_to_all_single(send, recv)      # combine
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        per_rank_bytes = send.numel() * 2
        print(json.dumps({"what": "NCCL all_to_all_single, reference EP exchange volume, collectives only",
                          "world": world, "chunks_per_step": n_chunks,
                          "bytes_per_rank_per_direction_per_chunk": per_rank_bytes,
                          "ms_per_step": float(t.item()), "ms_per_chunk": float(t.item()) / n_chunks,
                          "ours_bytes_per_rank_per_chunk": {"records_out": chunk * 48, "rows_back": chunk * 16}}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
