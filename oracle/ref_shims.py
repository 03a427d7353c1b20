"""TEST INFRASTRUCTURE ONLY -- import shims that let the *unmodified* reference
(`/root/reference/switch_nerf`) run on CPU in this container.

The reference hard-imports two packages that are not vendored and cannot be
installed offline:

  * microsoft/tutel @ 56dbd664341cf6485c9fa292955f77d3ac918a65 (install_tutel.md:3)
  * timm (only `timm.models.layers.trunc_normal_`)

The shims restate the published semantics of the handful of Tutel entry points
the hot path touches, anchored on the reference's own call sites:

  tutel.jit_kernels.gating.fast_cumsum_sub_one   -> cumsum(mask, 0) - 1
        (call sites tutel_fast_dispatch.py:138,190; CPU fallback named
         torch_cumsum_sub_one at :11)
  tutel.jit_kernels.sparse.create_forward        -> out[idx*cap+loc] += gate*x   if idx>=0 and loc<cap
  tutel.jit_kernels.sparse.create_backward_data  -> y[s] = gate*buf[idx*cap+loc] else 0
  tutel.jit_kernels.sparse.create_backward_gate  -> g[s] = <buf[row], x[s]> else 0
        (call sites tutel_fast_dispatch.py:27,36,43,61,70,76; in-tree analogue
         with identical addressing: tutel_sparse_nobatch.py:24-34,45-63,74-133)
  tutel.impls.communicate.{get_world_size,get_world_rank,all_to_all_single,
        simple_all_reduce,create_groups_from_world}  -> single-rank identities

This module is used by `oracle/make_golden.py` (which writes tests/golden/*)
and by the CPU tests that pin `oracle/switch_nerf_oracle.py` against the real
reference when `/root/reference` exists.  Nothing here is imported by the
product package `switch_nerf_b200`.

Because the K4/K5/K7 Tutel kernels themselves are not in the tree, parity at
that boundary is anchored on the in-tree K1/K2/K3 sources and call sites
("parity unpinned" for the external Tutel binary itself).
"""
import sys
import types
import contextlib
from argparse import Namespace

import torch

from oracle.install_ref import reference_root

# /root/reference in the build container, the verbatim copy under baseline/_ref on the GPU box (oracle/install_ref.py)
REFERENCE_ROOT = reference_root() or "/root/reference"


def _sparse_module():
    m = types.ModuleType("tutel.jit_kernels.sparse")

    def create_forward(dtype, is_cuda=False):
        def func_fwd(gates, indices, locations, x, out, extra):
            samples, hidden, capacity = extra
            keep = (indices >= 0) & (locations < capacity) & (locations >= 0)
            rows = (indices.long() * capacity + locations.long())[keep]
            g = gates[:, 0] if gates.dim() == 2 else gates
            out.index_add_(0, rows, x[keep] * g[: samples][keep].unsqueeze(1).to(x.dtype))
        return func_fwd

    def create_backward_data(dtype, is_cuda=False):
        def func_bwd_data(gates, indices, locations, y, buf, extra):
            samples, hidden, capacity = extra
            keep = (indices >= 0) & (locations < capacity) & (locations >= 0)
            rows = (indices.long() * capacity + locations.long()).clamp(0, buf.shape[0] - 1)
            g = gates[:, 0] if gates.dim() == 2 else gates
            vals = buf[rows] * g[: samples].unsqueeze(1).to(buf.dtype)
            y.copy_(torch.where(keep.unsqueeze(1), vals, torch.zeros_like(vals)))
        return func_bwd_data

    def create_backward_gate(dtype, is_cuda=False):
        def func_bwd_gate(grad_gates, indices, locations, x, buf, extra):
            samples, hidden, capacity = extra
            keep = (indices >= 0) & (locations < capacity) & (locations >= 0)
            rows = (indices.long() * capacity + locations.long()).clamp(0, buf.shape[0] - 1)
            dots = (buf[rows] * x).sum(1)
            grad_gates.copy_(torch.where(keep, dots, torch.zeros_like(dots)))
        return func_bwd_gate

    m.create_forward = create_forward
    m.create_backward_data = create_backward_data
    m.create_backward_gate = create_backward_gate
    return m


def install_shims():
    """Register stub modules for tutel.* and timm.* (idempotent)."""
    if "tutel" in sys.modules and getattr(sys.modules["tutel"], "_snb_shim", False):
        return

    tutel = types.ModuleType("tutel")
    tutel._snb_shim = True
    impls = types.ModuleType("tutel.impls")
    comm = types.ModuleType("tutel.impls.communicate")
    jitc = types.ModuleType("tutel.impls.jit_compiler")
    jk = types.ModuleType("tutel.jit_kernels")
    gating = types.ModuleType("tutel.jit_kernels.gating")
    sparse = _sparse_module()

    comm.get_world_size = lambda group=None: 1
    comm.get_world_rank = lambda group=None: 0
    comm.all_to_all_single = lambda t, group=None, **kw: t
    comm.simple_all_reduce = lambda t, group=None, op=None: t
    comm.TUTEL_GROUPING_CACHE = {}

    def create_groups_from_world(group_count=1, include_init=None):
        return Namespace(data_group=None, model_group=None, global_size=1, global_rank=0,
                         group_count=group_count, is_distributed=False,
                         local_device=torch.device("cpu"), local_rank=0, dist_print=print)
    comm.create_groups_from_world = create_groups_from_world

    class _JitCompiler:
        @staticmethod
        def generate_kernel(keyword_dict, template):
            """CUDA tensors: the real Tutel compiles the reference's kernel strings (tutel_sparse_nobatch.py:21-35,
            42-64, 71-134) with NVRTC; the shim recognises which of the three it was handed and returns the same
            torch restatement `generate_cpu_kernel` gives (the ops run on whatever device the tensors live on)."""
            if "atomicAdd" in template:
                return _JitCompiler.generate_cpu_kernel(0)
            if "grad_gates1_s" in template:
                return _JitCompiler.generate_cpu_kernel(2)
            return _JitCompiler.generate_cpu_kernel(1)

        @staticmethod
        def generate_cpu_kernel(kernel_type):
            """CPU stand-ins for the three CUDA-string kernels of
            tutel_sparse_nobatch.py (K1 :21-35, K2 :42-64, K3 :71-134); same
            argument order as the `execute` signatures there."""
            def k_fwd(gates, indices, locations, begin, x, out, extra):
                samples, hidden, _ = extra
                keep = indices >= 0
                rows = (begin.long()[indices.long().clamp_min(0)] + locations.long())[keep]
                g = gates[:, 0] if gates.dim() == 2 else gates
                out.index_add_(0, rows, x[keep] * g[:samples][keep].unsqueeze(1).to(x.dtype))

            def k_bwd_data(gates, indices, locations, begin, y, buf, extra):
                samples, hidden, _ = extra
                keep = indices >= 0
                rows = (begin.long()[indices.long().clamp_min(0)] + locations.long()).clamp(0, buf.shape[0] - 1)
                g = gates[:, 0] if gates.dim() == 2 else gates
                vals = buf[rows] * g[:samples].unsqueeze(1).to(buf.dtype)
                y.copy_(torch.where(keep.unsqueeze(1), vals, torch.zeros_like(vals)))

            def k_bwd_gate(grad_gates, indices, locations, begin, x, buf, extra):
                keep = indices >= 0
                rows = (begin.long()[indices.long().clamp_min(0)] + locations.long()).clamp(0, buf.shape[0] - 1)
                dots = (buf[rows] * x).sum(1)
                grad_gates.copy_(torch.where(keep, dots, torch.zeros_like(dots)))
            return [k_fwd, k_bwd_data, k_bwd_gate][kernel_type]
    jitc.IS_HIP_EXTENSION = False
    jitc.JitCompiler = _JitCompiler

    gating.fast_cumsum_sub_one = lambda m, dim=0: torch.cumsum(m, dim) - 1
    gating.torch_cumsum_sub_one = gating.fast_cumsum_sub_one

    tutel.impls, tutel.jit_kernels, tutel.net = impls, jk, comm
    impls.communicate, impls.jit_compiler = comm, jitc
    jk.gating, jk.sparse = gating, sparse

    for name, mod in [("tutel", tutel), ("tutel.impls", impls), ("tutel.impls.communicate", comm),
                      ("tutel.impls.jit_compiler", jitc), ("tutel.net", comm),
                      ("tutel.jit_kernels", jk), ("tutel.jit_kernels.gating", gating),
                      ("tutel.jit_kernels.sparse", sparse)]:
        sys.modules[name] = mod

    timm = types.ModuleType("timm")
    timm_models = types.ModuleType("timm.models")
    timm_layers = types.ModuleType("timm.models.layers")
    timm_layers.trunc_normal_ = torch.nn.init.trunc_normal_
    timm.models, timm_models.layers = timm_models, timm_layers
    sys.modules.update({"timm": timm, "timm.models": timm_models, "timm.models.layers": timm_layers})

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


@contextlib.contextmanager
def stable_argsort():
    """F9 / Appendix A.4: the reference calls `argsort` without stable=True inside
    compute_sorted_location (tutel_fast_dispatch.py:136-139); tie order is then
    implementation-defined.  The parity contract fixes ties to ascending sample
    index, so the reference is run with argsort forced stable."""
    orig = torch.Tensor.argsort

    def _stable(self, *args, **kwargs):
        kwargs.setdefault("stable", True)
        if "dim" not in kwargs and len(args) >= 1:
            kwargs["dim"] = args[0]
            args = args[1:]
        return orig(self, *args, **kwargs)

    torch.Tensor.argsort = _stable
    try:
        yield
    finally:
        torch.Tensor.argsort = orig


from switch_nerf_b200.configs import building_model_cfg, make_hparams  # noqa: E402,F401


def build_reference_model(hparams, appearance_count=16, xyz_dim=3, seed=0):
    """Instantiate the unmodified reference model (models/nerf_moe.py:1004-1041)."""
    install_shims()
    from switch_nerf.models.nerf_moe import get_nerf_moe_inner
    torch.manual_seed(seed)
    model = get_nerf_moe_inner(hparams, appearance_count, xyz_dim)
    return model
