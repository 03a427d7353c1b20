"""Host-side mirror of `switch_nerf.models.nerf_moe` (reference models/nerf_moe.py:103-455,
1004-1041) over the C ABI.

`NeRFMoE` here is a parameter container with the reference's exact module tree -- so
`state_dict()` keys/shapes, `load_state_dict` of reference checkpoints,
optimizers and DDP wrappers work unchanged -- whose `forward` hands the whole chunk to
`snb_moe_forward` (one call = positional encoding -> external gate -> top-1 routing -> expert
MLPs -> sigma/RGB heads).  No torch op touches the activations.

Construction consumes the torch RNG in the same order as the reference (embedding_a, then the
layer tags [0, 1, 2, xyz, sigma, color, moe_external_gate, gate_input_norm]; experts seeded by
`seeds=(1, rank+1, 1)`, nerf_moe.py:284, tutel_moe_layer_nobatch.py:636-703), so the same
`torch.manual_seed` gives bit-identical initial weights.

Scope: forward on the fused kernels; the backward for the parameters is attached by rendering.render_rays
(csrc/snb_backward.cu, fp32 kernels).
"""
import ctypes as C
import weakref
from typing import Optional

import torch
from torch import nn

from . import _lib as L


class Mlp(nn.Module):
    """Parameter layout of reference Mlp (nerf_moe.py:16-28): `fcs.{i}.weight/bias`."""

    def __init__(self, in_features, hidden_features, out_features, layer_num, skips=None):
        super().__init__()
        self.layer_num, self.skips = layer_num, skips
        self.fcs = nn.ModuleList()
        for i in range(layer_num):
            in_ch = in_features if i == 0 else hidden_features
            out_ch = out_features if i == layer_num - 1 else hidden_features
            self.fcs.append(nn.Linear(in_ch, out_ch))


class ExpertMLP(nn.Module):
    """Parameter layout of reference ExpertMLP (tutel_moe_layer_nobatch.py:837-866):
    `weights.{j}` [E, in, out], `bias.{j}` [E, 1, out], each expert initialised from a fresh
    nn.Linear (transposed) times init_factor."""

    def __init__(self, model_dim, local_experts, layer_num, skips=None, init_factor=1.0):
        super().__init__()
        self.model_dim, self.local_experts, self.layer_num, self.skips = model_dim, local_experts, layer_num, skips
        self.weights, self.bias = nn.ParameterList(), nn.ParameterList()
        for _ in range(layer_num):
            w = nn.Parameter(torch.zeros(local_experts, model_dim, model_dim))
            b = nn.Parameter(torch.zeros(local_experts, 1, model_dim))
            for i in range(local_experts):
                fc = nn.Linear(model_dim, model_dim)
                with torch.no_grad():
                    w[i, :, :], b[i, :, :] = fc.weight.t() * init_factor, fc.bias * init_factor
            self.weights.append(w)
            self.bias.append(b)


class SingleExpert(nn.Module):
    """Parameter layout of reference SingleExpert (tutel_moe_layer_nobatch.py:927-985), the per-expert module of the
    `seqexperts` checkpoint layout: `layers.{j}.weight [out, in]`, `layers.{j}.bias [out]`.  A container only: the fused
    path packs the stacked `expertmlp` layout (checkpoint.to_expertmlp converts)."""

    def __init__(self, model_dim, layer_num, skips=None, init_factor=1.0, **_):
        super().__init__()
        self.model_dim, self.layer_num, self.skips = model_dim, layer_num, skips
        self.layers = nn.ModuleList()
        for _i in range(layer_num):
            fc = nn.Linear(model_dim, model_dim)
            if init_factor != 1.0:
                with torch.no_grad():
                    fc.weight.multiply_(init_factor)
                    fc.bias.multiply_(init_factor)
            self.layers.append(fc)


def fast_cumsum_sub_one(mask: torch.Tensor, dim: int = 0) -> torch.Tensor:
    """Tutel `jit_kernels.gating.fast_cumsum_sub_one` (re-exported by tutel_moe_nobatch.py:6; call sites
    tutel_fast_dispatch.py:138, 190): inclusive cumsum over samples minus one, for a one-hot [S, E] CUDA mask.
    Computed by the routing kernel: for a one-hot row the only entry callers read is locations[s, idx[s]]
    (tutel_fast_dispatch.py:192-194), which is `loc[s]` of snb_route_top1 in sample order; the other entries are the
    running counts of the other experts, rebuilt here from the same per-expert locations."""
    if dim != 0 or mask.dim() != 2 or not mask.is_cuda:
        raise L.SnbError("fast_cumsum_sub_one: expects a CUDA [S, E] mask and dim=0 (no CPU path)")
    S, E = mask.shape
    gates = mask.to(torch.float32).contiguous()            # one-hot rows: argmax == the hot expert
    dev = mask.device
    idx = torch.empty(S, dtype=torch.int32, device=dev)
    loc = torch.empty(S, dtype=torch.int32, device=dev)
    gv = torch.empty(S, dtype=torch.float32, device=dev)
    counts = torch.empty(E, dtype=torch.int32, device=dev)
    cap = torch.empty(1, dtype=torch.int32, device=dev)
    l_aux = torch.empty(1, dtype=torch.float32, device=dev)
    lib = L.lib()
    with torch.cuda.device(dev):
        nb = lib.snb_route_workspace_bytes(S, E)
        ws = L.Workspace.get(nb, dev)
        L.check(lib.snb_route_top1(L.ptr(gates), S, E, 1.0, 0, L.ptr(idx), L.ptr(loc), L.ptr(gv), L.ptr(counts), L.ptr(cap),
                                   L.ptr(l_aux), L.ptr(ws), nb, L.stream_handle()))
    # running count of every expert at every sample: scatter the hot entries, then carry them forward
    out = torch.full((S, E), -1, dtype=mask.dtype, device=dev)
    hot = mask.sum(1) > 0                                   # all-zero rows route nowhere
    rows = torch.nonzero(hot).view(-1)
    out[rows, idx.long()[rows]] = loc.to(mask.dtype)[rows]
    return torch.cummax(out, 0).values


class TopKGate(nn.Module):
    """Parameter layout of reference TopKGate (tutel_moe_layer_nobatch.py:29-96): `wg.weight` [E, gate_dim]."""

    def __init__(self, gate_dim, num_global_experts, capacity_factor, batch_prioritized_routing):
        super().__init__()
        self.wg = nn.Linear(gate_dim, num_global_experts, bias=False)
        self.capacity_factor = float(capacity_factor)
        self.batch_prioritized_routing = bool(batch_prioritized_routing)
        self.top_k = 1


# NeRFMoE instances by id (weak): lets MOELayer.forward find the model handle without a reference cycle in the module tree
_OWNERS = weakref.WeakValueDictionary()


class MOELayer(nn.Module):
    """Parameter layout of reference MOELayer (tutel_moe_layer_nobatch.py:428-731):
    `experts.0.{weights,bias}.{j}`, `gates.0.wg.weight`; `moe_no_batch` toggled by
    Runner.set_no_batch (runner.py:947-956)."""

    def __init__(self, gate_type, model_dim, experts, seeds=None, moe_no_batch=False, return_gates=False,
                 ep_world=1, **_):
        super().__init__()
        assert gate_type["type"] == "top" and gate_type["k"] == 1, "hot path = top-1 switch routing"
        assert experts["type"] == "expertmlp", "expert type of the training configs (README.md:70)"
        # MOELayer.global_expert_count (tutel_moe_layer_nobatch.py:480): local experts x ranks of the MoE group
        self.num_local_experts = experts["count_per_node"]
        self.num_global_experts = self.num_local_experts * int(ep_world)
        self.model_dim, self.moe_no_batch, self.return_gates = model_dim, moe_no_batch, return_gates
        if seeds is not None and seeds[1] is not None:
            torch.manual_seed(seeds[1])
        self.experts = nn.ModuleList([ExpertMLP(model_dim, self.num_local_experts, experts["layer_num"],
                                                experts["skips"], experts["init_factor"])])
        if seeds is not None and seeds[0] is not None:
            torch.manual_seed(seeds[0])
        self.gates = nn.ModuleList([TopKGate(gate_type.get("gate_dim", model_dim), self.num_global_experts,
                                             gate_type["capacity_factor"], gate_type["batch_prioritized_routing"])])
        if seeds is not None and len(seeds) > 2 and seeds[2] is not None:
            torch.manual_seed(seeds[2])
        if int(ep_world) > 1:
            # scan_expert_func of the reference (nerf_moe.py:139, 284; tutel_moe_layer_nobatch.py:675-678): sharded
            # expert parameters differ per rank, DDP must neither broadcast nor all-reduce them
            for p in self.experts.parameters():
                setattr(p, "skip_allreduce", True)
        self.l_aux = None
        self.gate_extras = None
        self._owner_key = None      # key of the NeRFMoE that owns the packed weights in _OWNERS (set by NeRFMoE.__init__)

    def _find_owner(self):
        o = _OWNERS.get(self._owner_key) if self._owner_key is not None else None
        if o is not None and any(l is self for l in o.layers.values()):
            return o
        for o in list(_OWNERS.values()):           # e.g. after copy.deepcopy of the model
            if any(l is self for l in o.layers.values()):
                self._owner_key = id(o)
                return o
        return None

    def forward(self, input: torch.Tensor, gate_index=0, **kwargs):
        """reference MOELayer.forward (tutel_moe_layer_nobatch.py:733-797): top-1 capacity routing on
        softmax(gate_input @ wg^T), expert stack, combine.  Returns the result reshaped like `input`, carrying
        `.l_aux` and (if return_gates) `.gate_extras` as attributes, exactly like the reference.  Runs the fp32 CUDA
        path through snb_moe_layer_forward; NeRFMoE.forward itself uses the fused kernels and never calls this."""
        owner = self._find_owner()
        if owner is None:
            raise L.SnbError("MOELayer.forward needs the owning NeRFMoE (packed weights live in its model handle)")
        if gate_index != 0:
            raise NotImplementedError("the hot path has one gate per MoE layer")
        gate_input = kwargs.get("gate_input")
        original_shape, original_dtype = input.shape, input.dtype
        assert len(input.shape) >= 2, "Input data must be at least 2D tensor: (s)amples, .., (m)odel_dim"
        x = L.require_cuda_f32(input.reshape(-1, input.shape[-1]), "input", self.model_dim)
        g = None if gate_input is None else L.require_cuda_f32(gate_input.reshape(-1, gate_input.shape[-1]), "gate_input",
                                                               self.model_dim)
        S = x.shape[0]
        h = owner.handle()
        lib = L.lib()
        gate = self.gates[0]
        opts = L.RouteOpts(gate.capacity_factor, int(gate.batch_prioritized_routing), int(self.moe_no_batch))
        y = torch.empty(S, self.model_dim, dtype=torch.float32, device=x.device)
        idx = torch.empty(S, dtype=torch.int32, device=x.device)
        l_aux = torch.zeros(1, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            nbytes = lib.snb_workspace_bytes(h, S, opts.capacity_factor)
            ws = L.Workspace.get(nbytes, x.device)
            L.check(lib.snb_moe_layer_forward(h, L.ptr(x), L.ptr(g), S, C.byref(opts), L.ptr(y), L.ptr(idx), L.ptr(l_aux),
                                              L.ptr(ws), ws.numel(), L.stream_handle()))
        out = y.view(original_shape).to(original_dtype)
        self.l_aux = out.l_aux = l_aux[0]
        if self.return_gates:
            self.gate_extras = out.gate_extras = {"gates": idx.long().view(-1, 1)}     # topk(gates, 1).indices (:227-229)
        return out


moe_layer = MOELayer


def expert_parallel_env(args):
    """(rank, world) of the MoE process group.  The reference passes `group=args.single_data_group`
    (nerf_moe.py:289): a 1-rank group when `no_expert_parallel` (always true upstream, opts.py:125), else WORLD."""
    if getattr(args, "no_expert_parallel", True):
        return 0, 1
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(), dist.get_world_size()


def gather_expert_shards(t: torch.Tensor, group=None) -> torch.Tensor:
    """[E_local, ...] on every rank -> [E, ...] in rank order (expert e lives on rank e // E_local)."""
    import torch.distributed as dist
    t = t.contiguous()
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, t, group=group)
    return torch.cat(parts, 0)


class NeRFMoE(nn.Module):
    """Drop-in for reference NeRFMoE / MipNeRFMoE (nerf_moe.py:103-455 / 458-810), Building /
    Mission-Bay topology: xyz -> external gate -> MoE layer "0" -> sigma, layer "1" -> dir/appearance
    concat -> layer "2" -> colour."""

    def __init__(self, args, pos_xyz_dim: int, pos_dir_dim: int, appearance_dim: int, affine_appearance: bool,
                 appearance_count: int, rgb_dim: int, xyz_dim: int, sigma_activation=None, mip: bool = False):
        super().__init__()
        cfg = args.layer_cfg
        lay = cfg["layers"]
        self._check_topology(args, cfg, rgb_dim, affine_appearance, pos_dir_dim, appearance_dim)
        self.args, self.layer_cfg, self.mip = args, cfg, mip
        self.xyz_dim, self.pos_xyz_dim, self.pos_dir_dim = xyz_dim, pos_xyz_dim, pos_dir_dim
        self.appearance_dim, self.appearance_count = appearance_dim, appearance_count
        self._ddp_params_and_buffers_to_ignore = []
        self._ep_rank, self._ep_world = 0, 1
        self.precision = "bf16" if getattr(args, "amp_use_bfloat16", False) else "fp32"
        self.embedding_a = nn.Embedding(appearance_count, appearance_dim)
        self.layers = nn.ModuleDict()
        tags = [str(i) for i in range(cfg["layer_num_main"])] + ["xyz", "sigma", "color", "moe_external_gate",
                                                                  "gate_input_norm"]
        for tag in tags:
            c = lay[tag]
            if c["type"] == "mlp":
                self.layers[tag] = Mlp(c["in_ch"], c["h_ch"], c["out_ch"], c["num"], c.get("skips"))
            elif c["type"] == "moe":
                gate_type = {"type": c["gate_type"], "k": c["k"], "capacity_factor": args.moe_capacity_factor,
                             "batch_prioritized_routing": args.batch_prioritized_routing,
                             "gate_dim": c.get("gate_dim", c["in_ch"])}
                experts = {"type": getattr(args, "moe_expert_type", "expertmlp"),
                           "count_per_node": c.get("local_expert_num") or args.moe_local_expert_num,
                           "layer_num": c["num"], "skips": c["skips"], "init_factor": c["init_factor"]}
                rank = getattr(getattr(args, "parallel_env", None), "global_rank", 0)
                self._ep_rank, self._ep_world = expert_parallel_env(args)
                self.layers[tag] = moe_layer(gate_type=gate_type, model_dim=c["in_ch"], experts=experts,
                                             seeds=(1, rank + 1, 1), moe_no_batch=False,
                                             return_gates=getattr(args, "moe_return_gates", False),
                                             ep_world=self._ep_world)
                self.layers[tag]._owner_key = id(self)
                _OWNERS[id(self)] = self
            elif c["type"] == "layernorm":
                self.layers[tag] = nn.LayerNorm(c["in_ch"])
        self._handle = None
        self._packed_versions = None
        self._packed_device = None
        self._ep_group = None      # ExpertParallelGroup once attached (expert_parallel.py)
        # checkpoints in the seqexperts layout / with DDP's `module.` prefix load as they are (checkpoint.py)
        self._register_load_state_dict_pre_hook(self._normalise_state_dict)

    @staticmethod
    def _normalise_state_dict(state_dict, prefix, *_):
        if prefix:
            return
        from .checkpoint import to_expertmlp
        if any(k.startswith("module.") or ".experts.0.experts." in k for k in state_dict):
            fixed = to_expertmlp(state_dict)
            state_dict.clear()
            state_dict.update(fixed)

    # copies / pickles never share the C handle (it is re-created on first use of the copy)
    def __getstate__(self):
        st = self.__dict__.copy()
        st["_handle"], st["_packed_versions"], st["_packed_device"], st["_ep_group"] = None, None, None, None
        return st

    def __setstate__(self, st):
        super().__setstate__(st)
        _OWNERS[id(self)] = self

    # -- reference API -----------------------------------------------------------------
    @staticmethod
    def _check_topology(args, cfg, rgb_dim, affine_appearance, pos_dir_dim, appearance_dim):
        lay = cfg["layers"]
        ok = (cfg["layer_num_main"] == 3 and str(cfg["sigma_tag"]) == "0" and str(cfg["dir_tag"]) == "1"
              and str(cfg["color_tag"]) == "2" and lay["0"]["type"] == "moe" and rgb_dim == 3
              and not affine_appearance and pos_dir_dim > 0 and appearance_dim > 0
              and getattr(args, "use_moe_external_gate", False) and getattr(args, "use_gate_input_norm", False)
              and lay["xyz"].get("act", "none") == "none" and lay["0"].get("act") == "relu"
              and lay["1"].get("act", "none") == "none" and lay["2"].get("act") == "relu"
              and lay["moe_external_gate"].get("act", "none") == "none" and lay["0"]["k"] == 1
              and getattr(args, "shifted_softplus", True))
        if not ok:
            raise NotImplementedError(
                "switch_nerf_b200 implements the Switch-NeRF hot path: the building.yaml / mission_bay.yaml "
                "topology (xyz -> external gate + LayerNorm -> one top-1 MoE layer -> sigma / dir+appearance / colour)")

    def add_param_to_skip_allreduce(self, param_name):
        self._ddp_params_and_buffers_to_ignore.append(param_name)

    def set_no_batch(self, mode=True):
        for net in self.modules():
            if type(net) == MOELayer:
                net.moe_no_batch = mode

    # -- packing -----------------------------------------------------------------------
    def _weights_struct(self):
        lay, keep = self.layers, []

        def p(t):
            t = t.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            keep.append(t)
            return t.data_ptr()

        w = L.Weights()
        w.xyz_w, w.xyz_b = p(lay["xyz"].fcs[0].weight), p(lay["xyz"].fcs[0].bias)
        for i, fc in enumerate(lay["moe_external_gate"].fcs):
            w.gate_w[i], w.gate_b[i] = p(fc.weight), p(fc.bias)
        w.ln_w, w.ln_b = p(lay["gate_input_norm"].weight), p(lay["gate_input_norm"].bias)
        moe = lay["0"]
        w.wg = p(moe.gates[0].wg.weight)
        for j in range(moe.experts[0].layer_num):
            ew, eb = moe.experts[0].weights[j], moe.experts[0].bias[j]
            if self._ep_world > 1:      # parameters are sharded E/W per rank: launch #2 packs the full set
                ew, eb = gather_expert_shards(ew.detach()), gather_expert_shards(eb.detach())
            w.expert_w[j], w.expert_b[j] = p(ew), p(eb)
        w.l1_w, w.l1_b = p(lay["1"].fcs[0].weight), p(lay["1"].fcs[0].bias)
        w.l2_w, w.l2_b = p(lay["2"].fcs[0].weight), p(lay["2"].fcs[0].bias)
        w.sigma_w, w.sigma_b = p(lay["sigma"].fcs[0].weight), p(lay["sigma"].fcs[0].bias)
        w.color_w, w.color_b = p(lay["color"].fcs[0].weight), p(lay["color"].fcs[0].bias)
        w.emb_a = p(self.embedding_a.weight)
        return w, keep

    def _grad_params(self):
        """The parameters in the field order of snb_weights / snb_grads: [(field, index or None, parameter)]."""
        lay, moe = self.layers, self.layers["0"]
        out = [("xyz_w", None, lay["xyz"].fcs[0].weight), ("xyz_b", None, lay["xyz"].fcs[0].bias)]
        for i, fc in enumerate(lay["moe_external_gate"].fcs):
            out += [("gate_w", i, fc.weight), ("gate_b", i, fc.bias)]
        out += [("ln_w", None, lay["gate_input_norm"].weight), ("ln_b", None, lay["gate_input_norm"].bias),
                ("wg", None, moe.gates[0].wg.weight)]
        for j in range(moe.experts[0].layer_num):
            out += [("expert_w", j, moe.experts[0].weights[j]), ("expert_b", j, moe.experts[0].bias[j])]
        out += [("l1_w", None, lay["1"].fcs[0].weight), ("l1_b", None, lay["1"].fcs[0].bias),
                ("l2_w", None, lay["2"].fcs[0].weight), ("l2_b", None, lay["2"].fcs[0].bias),
                ("sigma_w", None, lay["sigma"].fcs[0].weight), ("sigma_b", None, lay["sigma"].fcs[0].bias),
                ("color_w", None, lay["color"].fcs[0].weight), ("color_b", None, lay["color"].fcs[0].bias),
                ("emb_a", None, self.embedding_a.weight)]
        return out

    def _desc(self):
        lay = self.layer_cfg["layers"]
        moe = self.layers["0"]
        d = L.ModelDesc()
        d.num_experts, d.width = moe.num_global_experts, lay["0"]["in_ch"]
        d.expert_layers = lay["0"]["num"]
        d.skip_layer = lay["0"]["skips"][0] if lay["0"].get("skips") else -1
        d.gate_layers = lay["moe_external_gate"]["num"]
        d.pos_xyz_freqs, d.pos_dir_freqs = self.pos_xyz_dim, self.pos_dir_dim
        d.appearance_dim, d.appearance_count = self.appearance_dim, self.appearance_count
        d.hidden2, d.mip = lay["2"]["out_ch"], int(self.mip)
        return d

    def handle(self):
        """snb_model_t* for the current parameter values (re-packed when any parameter changed)."""
        params = list(self.parameters())
        dev = params[0].device
        if dev.type != "cuda":
            raise L.SnbError("NeRFMoE parameters must live on a CUDA device (no CPU path)")
        versions = tuple(p._version for p in params) + tuple(p.data_ptr() for p in params)
        lib = L.lib()
        with torch.cuda.device(dev):
            if self._handle is None or self._packed_device != dev:
                self.release()
                w, keep = self._weights_struct()
                h = C.c_void_p()
                L.check(lib.snb_model_create(C.byref(self._desc()), C.byref(w), L.stream_handle(), C.byref(h)))
                self._handle, self._packed_versions, self._packed_device = h, versions, dev
                if self._ep_group is None and self._ep_world > 1:
                    from .expert_parallel import ExpertParallelGroup
                    self._ep_group = ExpertParallelGroup(
                        self.layers["0"].num_global_experts, int(getattr(self.args, "model_chunk_size", 131072)),
                        float(self.layers["0"].gates[0].capacity_factor), rank=self._ep_rank, world=self._ep_world,
                        device=dev)
                    self._ep_group._models.append(self)
                if self._ep_group is not None:
                    L.check(lib.snb_model_attach_a2a(h, self._ep_group._h))
            elif versions != self._packed_versions:
                w, keep = self._weights_struct()
                L.check(lib.snb_model_update(self._handle, C.byref(w), L.stream_handle()))
                self._packed_versions = versions
        return self._handle

    def mark_dirty(self):
        """Force a re-pack of the device copies at the next call.  `handle()` notices parameter changes through
        `Parameter._version` (optimizer steps, `load_state_dict`, in-place ops on the parameter); writes through
        `param.data` (`p.data.copy_()`, common in EMA / manual loading code) do not bump it -- call this after them.
        With sharded experts the re-pack gathers the shards (a collective): call it on every rank."""
        self._packed_versions = None

    def tuning(self, **changes):
        """Read (and optionally change) the kernel-selection / pipeline knobs of this model (snb_tuning: route_sms,
        pipe_depth, ts, cta_group_*, route_full, no_overlap, ...): `model.tuning(route_sms=16)`; returns the current values."""
        h = self.handle()
        t = L.Tuning()
        L.check(L.lib().snb_model_get_tuning(h, C.byref(t)))
        if changes:
            for k, v in changes.items():
                if not hasattr(t, k):
                    raise AttributeError(f"snb_tuning has no field {k!r}")
                setattr(t, k, int(v))
            L.check(L.lib().snb_model_set_tuning(h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in L.Tuning._fields_}

    def release(self):
        if self._handle is not None:
            L.lib().snb_model_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def route_opts(self):
        gate = self.layers["0"].gates[0]
        return L.RouteOpts(gate.capacity_factor, int(gate.batch_prioritized_routing), int(self.layers["0"].moe_no_batch))

    # -- forward -----------------------------------------------------------------------
    def forward(self, x: torch.Tensor, sigma_only: bool = False, sigma_noise: Optional[torch.Tensor] = None,
                return_debug: bool = False):
        expected = self.xyz_dim + 3 + 1
        if x.dim() != 2 or x.shape[1] != expected:   # nerf_moe.py:322-328
            raise Exception('Unexpected input shape: {} (expected: {}, xyz_dim: {})'.format(x.shape, expected, self.xyz_dim))
        if sigma_only:
            raise NotImplementedError("sigma_only is never set by rendering.render_rays for this topology")
        x = L.require_cuda_f32(x, "x")
        S = x.shape[0]
        E = self.layers["0"].num_global_experts
        h = self.handle()
        lib = L.lib()
        opts = self.route_opts()
        out = torch.empty(S, 4, dtype=torch.float32, device=x.device)
        idx = torch.empty(S, dtype=torch.int32, device=x.device)
        l_aux = torch.zeros(1, dtype=torch.float32, device=x.device)
        dbg_g = torch.empty(S, E, dtype=torch.float32, device=x.device) if return_debug else None
        dbg_l = torch.empty(S, dtype=torch.int32, device=x.device) if return_debug else None
        noise = None if sigma_noise is None else L.require_cuda_f32(sigma_noise.reshape(-1), "sigma_noise")
        with torch.cuda.device(x.device):
            nbytes = lib.snb_workspace_bytes(h, S, opts.capacity_factor)
            ws = L.Workspace.get(nbytes, x.device)
            L.check(lib.snb_moe_forward(h, L.ptr(x), S, L.ptr(noise), C.byref(opts), L.PRECISIONS[self.precision],
                                        L.ptr(out), L.ptr(idx), L.ptr(l_aux), L.ptr(dbg_g), L.ptr(dbg_l), L.ptr(ws),
                                        ws.numel(), L.stream_handle()))
        extras = {"moe_loss": l_aux}                                   # torch.stack(moe_loss), nerf_moe.py:448-450
        if getattr(self.args, "moe_return_gates", False):
            extras["moe_gates"] = [idx.long().view(-1, 1)]             # topk indices int64 [S,1], :227-229
        if return_debug:
            extras["debug_gates"], extras["debug_loc"] = dbg_g, dbg_l
        return {"outputs": out, "extras": extras}


class MipNeRFMoE(NeRFMoE):
    """reference MipNeRFMoE (nerf_moe.py:458-810): same flow, MipEmbedder over [mean, cov_diag]."""

    def __init__(self, *a, **k):
        super().__init__(*a, mip=True, **k)
        self.xyz_dim = 6          # x = [mean(3), cov_diag(3), dir(3), image_index]
        # precision follows hparams.amp_use_bfloat16 like NeRFMoE: bf16 = the wide tcgen05 kernels (csrc/snb_tc_wide.cuh)


def get_nerf_moe_inner(hparams, appearance_count: int, xyz_dim: int, model_cfg_name="model") -> nn.Module:
    """reference get_nerf_moe_inner (nerf_moe.py:1004-1041)."""
    rgb_dim = 3 * ((hparams.sh_deg + 1) ** 2) if getattr(hparams, "sh_deg", None) is not None else 3
    model_cfg = getattr(hparams, model_cfg_name)
    hparams.layer_cfg = {k: model_cfg[k] for k in ("layer_num_main", "sigma_tag", "dir_tag", "color_tag", "layers")}
    name = getattr(hparams, "nerfmoe_class_name", "NeRFMoE")
    cls = MipNeRFMoE if name == "MipNeRFMoE" else NeRFMoE
    model = cls(hparams, hparams.pos_xyz_dim, hparams.pos_dir_dim, hparams.appearance_dim,
                hparams.affine_appearance, appearance_count, rgb_dim, xyz_dim, None)
    for name_, param in model.named_parameters():
        if hasattr(param, "skip_allreduce"):
            model.add_param_to_skip_allreduce(name_)
    return model
