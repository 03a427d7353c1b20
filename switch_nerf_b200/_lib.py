"""ctypes binding of libsnb.so -- the C ABI declared in include/switch_nerf_b200.h.

PyTorch is used by the callers only for device memory and streams; every call below passes raw
device pointers (`tensor.data_ptr()`), sizes and the current CUDA stream handle.  There is no
CPU path: `lib()` raises if the shared library is missing, and the library itself refuses to
create a model without a CUDA device.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsnb.so")

SNB_PREC_FP32 = 0
SNB_PREC_BF16 = 1
PRECISIONS = {"fp32": SNB_PREC_FP32, "bf16": SNB_PREC_BF16}


class SnbError(RuntimeError):
    pass


class ModelDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "num_experts", "width", "expert_layers", "skip_layer", "gate_layers", "pos_xyz_freqs",
        "pos_dir_freqs", "appearance_dim", "appearance_count", "hidden2", "mip")]


class Weights(C.Structure):
    _fields_ = [
        ("xyz_w", C.c_void_p), ("xyz_b", C.c_void_p),
        ("gate_w", C.c_void_p * 4), ("gate_b", C.c_void_p * 4),
        ("ln_w", C.c_void_p), ("ln_b", C.c_void_p), ("wg", C.c_void_p),
        ("expert_w", C.c_void_p * 16), ("expert_b", C.c_void_p * 16),
        ("l1_w", C.c_void_p), ("l1_b", C.c_void_p), ("l2_w", C.c_void_p), ("l2_b", C.c_void_p),
        ("sigma_w", C.c_void_p), ("sigma_b", C.c_void_p), ("color_w", C.c_void_p), ("color_b", C.c_void_p),
        ("emb_a", C.c_void_p)]


class RouteOpts(C.Structure):
    _fields_ = [("capacity_factor", C.c_double), ("bpr", C.c_int32), ("no_batch", C.c_int32)]


class RenderOpts(C.Structure):
    _fields_ = [("coarse_samples", C.c_int32), ("fine_samples", C.c_int32), ("model_chunk_size", C.c_int64),
                ("perturb", C.c_float), ("seed", C.c_uint64), ("white_bkgd", C.c_int32), ("precision", C.c_int32),
                ("route", RouteOpts), ("sigma_noise_coarse", C.c_void_p), ("sigma_noise_fine", C.c_void_p),
                ("resample_randomized", C.c_int32), ("last_delta_minus_zmax", C.c_int32)]


class BgDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("layers", "skip_layer", "width", "pos_xyz_freqs", "pos_dir_freqs", "appearance_dim",
                                         "appearance_count", "shifted_softplus")]


class BgWeights(C.Structure):
    _fields_ = [("w", C.c_void_p * 16), ("b", C.c_void_p * 16), ("final_w", C.c_void_p), ("final_b", C.c_void_p),
                ("dir_w", C.c_void_p), ("dir_b", C.c_void_p), ("sigma_w", C.c_void_p), ("sigma_b", C.c_void_p),
                ("rgb_w", C.c_void_p), ("rgb_b", C.c_void_p), ("emb_a", C.c_void_p)]


class Tuning(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "cta_group_front", "cta_group_back", "ts", "ts_front", "wide", "route_full", "no_overlap", "pipe_depth", "route_sms",
        "back_partition", "no_ray_source", "gather_h", "front_ab")]


class RenderOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "rgb", "depth", "depth_variance", "bg_lambda", "gate_loss_coarse", "gate_loss_fine",
        "moe_gates_coarse", "moe_gates_fine", "z_fine", "raw_coarse", "raw_fine", "rgb_coarse", "z_coarse")]


# name -> (restype, argtypes); mirrors include/switch_nerf_b200.h one to one
_SIGNATURES = {
    "snb_last_error": (C.c_char_p, []),
    "snb_version": (C.c_int, []),
    "snb_launch_count": (C.c_int64, []),
    "snb_profile_enable": (C.c_int, [C.c_int32]),
    "snb_profile_collect": (C.c_int, [C.POINTER(C.c_double)]),
    "snb_debug_timeline": (C.c_int, [C.POINTER(C.c_uint64), C.c_int32]),
    "snb_model_create": (C.c_int, [C.POINTER(ModelDesc), C.POINTER(Weights), C.c_void_p, C.POINTER(C.c_void_p)]),
    "snb_model_update": (C.c_int, [C.c_void_p, C.POINTER(Weights), C.c_void_p]),
    "snb_model_destroy": (None, [C.c_void_p]),
    "snb_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64, C.c_double]),
    "snb_a2a_init": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_double, C.POINTER(C.c_void_p)]),
    "snb_a2a_handle_bytes": (C.c_int32, []),
    "snb_a2a_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "snb_a2a_connect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "snb_a2a_connect_ptrs": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32]),
    "snb_a2a_local_base": (C.c_void_p, [C.c_void_p]),
    "snb_a2a_region_bytes": (C.c_size_t, [C.c_void_p]),
    "snb_model_attach_a2a": (C.c_int, [C.c_void_p, C.c_void_p]),
    "snb_a2a_disconnect": (C.c_int, [C.c_void_p]),
    "snb_a2a_finalize": (C.c_int, [C.c_void_p]),
    "snb_route_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "snb_route_top1": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_int32, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "snb_bg_create": (C.c_int, [C.POINTER(BgDesc), C.POINTER(BgWeights), C.c_void_p, C.POINTER(C.c_void_p)]),
    "snb_bg_update": (C.c_int, [C.c_void_p, C.POINTER(BgWeights), C.c_void_p]),
    "snb_bg_destroy": (None, [C.c_void_p]),
    "snb_bg_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64]),
    "snb_bg_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "snb_intersect_sphere": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "snb_depth2pts_outside": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "snb_model_get_tuning": (C.c_int, [C.c_void_p, C.POINTER(Tuning)]),
    "snb_model_set_tuning": (C.c_int, [C.c_void_p, C.POINTER(Tuning)]),
    "snb_moe_backward_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64, C.c_double]),
    "snb_moe_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(RouteOpts), C.c_void_p, C.c_void_p,
                                   C.POINTER(Weights), C.c_void_p, C.c_size_t, C.c_void_p]),
    "snb_composite_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    "snb_get_rays": (C.c_int, [C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_void_p,
                               C.c_float, C.c_float, C.POINTER(C.c_float), C.c_void_p, C.c_void_p]),
    "snb_route_select_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "snb_route_select": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_int32, C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                   C.c_void_p]),
    "snb_dispatch_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                   C.c_int64, C.c_void_p, C.c_void_p]),
    "snb_combine": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                              C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "snb_moe_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(RouteOpts), C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_size_t, C.c_void_p]),
    "snb_moe_layer_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(RouteOpts), C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "snb_render_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64, C.POINTER(RenderOpts)]),
    "snb_render_rays": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(RenderOpts),
                                  C.POINTER(RenderOut), C.c_void_p, C.c_size_t, C.c_void_p]),
    "snb_render_mip_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64, C.POINTER(RenderOpts)]),
    "snb_render_rays_mip": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.POINTER(RenderOpts), C.c_float, C.c_float, C.POINTER(RenderOut), C.c_void_p,
                                      C.c_size_t, C.c_void_p]),
    "snb_composite": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "snb_sample_pdf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                 C.c_void_p]),
    "snb_umma_microbench": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_uint64), C.c_void_p]),
    "snb_umma_selftest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def lib():
    """Load libsnb.so (built by `python -m switch_nerf_b200.build`).  Fails loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SnbError(f"{LIB_PATH} is missing: build it with `python -m switch_nerf_b200.build` "
                           "(there is no CPU or PyTorch path to run instead)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise SnbError(f"libsnb error {rc}: {lib().snb_last_error().decode()}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_handle():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda_f32(t, name, cols=None):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise SnbError(f"{name} must be a CUDA tensor (switch_nerf_b200 has no CPU path)")
    if t.dtype != torch.float32:
        t = t.float()
    if not t.is_contiguous():
        t = t.contiguous()
    if cols is not None and (t.dim() != 2 or t.shape[1] != cols):
        raise SnbError(f"{name}: expected shape [*, {cols}], got {tuple(t.shape)}")
    return t


class Workspace:
    """Caller-provided scratch for the C ABI, cached per device and grown on demand."""
    _cache = {}

    @classmethod
    def get(cls, nbytes, device, tag=None):
        key = (device.type, device.index, tag)
        buf = cls._cache.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = None
            cls._cache.pop(key, None)
            buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            cls._cache[key] = buf
        return buf
