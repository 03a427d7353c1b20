"""Expert-parallel group: the host-side plumbing of the P2P dispatch/combine (SURVEY §8e, configs[2]).

Mirrors what the reference gets from `tutel.net.create_groups_from_world` + the `group=` argument of
`moe_layer` (tutel_moe_layer_nobatch.py:443-460, 597-602; runner.py:100-101, 270, 381): experts are sharded
`moe_local_expert_num = E // world` per rank, rank r owns experts [r*E/W, (r+1)*E/W).

`torch.distributed` is used for exactly two things here -- exchanging the 64-byte CUDA IPC handles of the
symmetric buffers and one barrier after mapping them.  The data path is inside libsnb.so: st.global into peer
memory + release/acquire flags (csrc/snb_ep.cu); no collective is called per chunk.
"""
import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib as L


def owner_of_expert(expert: int, num_experts: int, world: int) -> int:
    """Rank that evaluates `expert` (contiguous blocks, as moe_local_expert_num implies)."""
    if num_experts % world:
        raise ValueError(f"{num_experts} experts do not shard evenly over {world} ranks")
    return expert // (num_experts // world)


def local_experts(rank: int, num_experts: int, world: int) -> range:
    n = num_experts // world
    if num_experts % world:
        raise ValueError(f"{num_experts} experts do not shard evenly over {world} ranks")
    return range(rank * n, (rank + 1) * n)


def gather_handles(handle: bytes, group=None, all_gather=None) -> bytes:
    """All-gather the per-rank handle blobs in rank order.  `all_gather(obj) -> list` can be injected (tests,
    MPI launchers); default = torch.distributed.all_gather_object on `group`."""
    if all_gather is None:
        import torch.distributed as dist

        def all_gather(obj):
            out = [None] * dist.get_world_size(group)
            dist.all_gather_object(out, obj, group=group)
            return out
    blobs = all_gather(bytes(handle))
    n = len(blobs[0])
    if any(len(b) != n for b in blobs):
        raise L.SnbError("expert-parallel handle exchange: ranks disagree on the handle size")
    return b"".join(blobs)


class ExpertParallelGroup:
    """One per process (= per GPU).  `attach(model)` switches a NeRFMoE to expert-parallel evaluation;
    all ranks must then make the same sequence of forward / render_rays calls (as with the reference's
    all_to_all)."""

    def __init__(self, num_experts: int, max_chunk_rows: int, max_capacity_factor: float = 1.0, group=None,
                 rank: Optional[int] = None, world: Optional[int] = None, device=None, connect: bool = True):
        import torch.distributed as dist
        if rank is None or world is None:
            if not dist.is_initialized():
                raise L.SnbError("ExpertParallelGroup needs torch.distributed (or explicit rank/world)")
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        self.rank, self.world, self.num_experts, self.group = int(rank), int(world), int(num_experts), group
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.local_experts = local_experts(self.rank, self.num_experts, self.world)
        self._h = C.c_void_p()
        self._models: List = []
        lib = L.lib()
        with torch.cuda.device(self.device):
            L.check(lib.snb_a2a_init(self.rank, self.world, self.num_experts, int(max_chunk_rows),
                                     float(max_capacity_factor), C.byref(self._h)))
        if connect:
            self.connect()

    # -- mapping the peers ---------------------------------------------------------------
    def export_handle(self) -> bytes:
        lib = L.lib()
        buf = C.create_string_buffer(lib.snb_a2a_handle_bytes())
        with torch.cuda.device(self.device):
            L.check(lib.snb_a2a_export(self._h, buf))
        return buf.raw

    def connect(self, all_gather=None):
        """Exchange IPC handles, map every peer's region, barrier."""
        lib = L.lib()
        if self.world > 1:
            blob = gather_handles(self.export_handle(), self.group, all_gather)
            with torch.cuda.device(self.device):
                L.check(lib.snb_a2a_connect(self._h, blob, len(blob)))
            if all_gather is None:
                import torch.distributed as dist
                torch.cuda.synchronize(self.device)
                dist.barrier(group=self.group)
            else:
                all_gather(b"mapped")
        return self

    def connect_ptrs(self, bases: Sequence[int]):
        """Same-process peers (peer access already enabled): base pointers in rank order."""
        arr = (C.c_void_p * self.world)(*[C.c_void_p(int(b)) for b in bases])
        with torch.cuda.device(self.device):
            L.check(L.lib().snb_a2a_connect_ptrs(self._h, arr, self.world))
        return self

    @property
    def local_base(self) -> int:
        return int(L.lib().snb_a2a_local_base(self._h) or 0)

    @property
    def region_bytes(self) -> int:
        return int(L.lib().snb_a2a_region_bytes(self._h))

    # -- models ----------------------------------------------------------------------------
    def attach(self, model):
        """model: switch_nerf_b200.nerf_moe.NeRFMoE on this group's device."""
        with torch.cuda.device(self.device):
            L.check(L.lib().snb_model_attach_a2a(model.handle(), self._h))
        model._ep_group = self
        self._models.append(model)
        return model

    def detach(self, model):
        if getattr(model, "_handle", None) is not None:
            with torch.cuda.device(self.device):
                L.check(L.lib().snb_model_attach_a2a(model._handle, None))
        model._ep_group = None
        if model in self._models:
            self._models.remove(model)

    def close(self, barrier=None):
        """Collective when world > 1: every rank unmaps its peers, then (after a barrier) frees its own region."""
        if self._h:
            for m in list(self._models):
                self.detach(m)
            lib = L.lib()
            with torch.cuda.device(self.device):
                lib.snb_a2a_disconnect(self._h)
            if self.world > 1:
                if barrier is not None:
                    barrier()
                else:
                    import torch.distributed as dist
                    if dist.is_available() and dist.is_initialized():
                        dist.barrier(group=self.group)
            with torch.cuda.device(self.device):
                lib.snb_a2a_finalize(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:      # interpreter teardown: no collective here, just release what this process owns
            if self._h:
                L.lib().snb_a2a_finalize(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass
