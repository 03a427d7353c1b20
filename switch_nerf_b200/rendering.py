"""Host-side mirror of `switch_nerf.rendering` (reference rendering.py:15-196) over the C ABI.

`render_rays` keeps the reference signature and `results` keys so `Runner._training_step` /
`Runner.render_image` (runner.py:1077-1123, 2835-2885) call it unchanged.  The whole ray pipeline
(coarse depths, point generation, chunked model evaluation, transmittance scan, inverse-CDF
resampling, fine pass, sorted merge and composite) is one `snb_render_rays` call; torch only
owns the output tensors.

Scope: foreground model, `bg_nerf=None`, `use_cascade=False`,
`pos_dir_dim > 0`, `sh_deg=None` -- the path of the Building-family configs without the
background sphere, which is rank-3 "next" work (8f).
"""
import ctypes as C
from argparse import Namespace
from typing import Dict, Optional, Tuple

import torch

from . import _lib as L
from .nerf_moe import NeRFMoE


_WARNED = False


def _warn_forward_only(model):
    """The fused path has no backward yet: say so once instead of letting loss.backward() fail later
    with an unrelated-looking autograd error."""
    global _WARNED
    if not _WARNED and model.training and torch.is_grad_enabled() and any(p.requires_grad for p in model.parameters()):
        import warnings
        warnings.warn("switch_nerf_b200.render_rays is forward-only: the results carry no autograd graph "
                      "(backward of the fused path is the next scope row)", RuntimeWarning, stacklevel=3)
        _WARNED = True


def _unwrap(nerf):
    return nerf.module if hasattr(nerf, "module") and isinstance(nerf.module, NeRFMoE) else nerf


def render_rays(nerf, bg_nerf, rays: torch.Tensor, image_indices: Optional[torch.Tensor], hparams: Namespace,
                sphere_center: Optional[torch.Tensor] = None, sphere_radius: Optional[torch.Tensor] = None,
                get_depth: bool = True, get_depth_variance: bool = True, get_bg_fg_rgb: bool = False,
                debug_taps: bool = False, seed: Optional[int] = None) -> Tuple[Dict[str, torch.Tensor], bool]:
    model = _unwrap(nerf)
    if not isinstance(model, NeRFMoE):
        raise L.SnbError("render_rays needs a switch_nerf_b200.nerf_moe.NeRFMoE model")
    if bg_nerf is not None:
        raise NotImplementedError("bg_nerf (background sphere NeRF) is outside the round-1 hot path (SURVEY 8f rank 3)")
    if getattr(hparams, "use_cascade", False):
        raise NotImplementedError("use_cascade is not used by any Switch-NeRF config")
    rays = L.require_cuda_f32(rays, "rays", cols=8)
    N = rays.shape[0]
    dev = rays.device
    Sc, Sf = int(hparams.coarse_samples), int(hparams.fine_samples)
    chunk = int(hparams.model_chunk_size)
    idx32 = None
    if image_indices is not None:
        idx32 = image_indices.to(device=dev, dtype=torch.int32).contiguous()
    perturb = float(hparams.perturb) if model.training else 0.0          # rendering.py:32
    _warn_forward_only(model)
    typ = "fine" if Sf > 0 else "coarse"

    opts = L.RenderOpts()
    opts.coarse_samples, opts.fine_samples, opts.model_chunk_size = Sc, Sf, chunk
    opts.perturb = perturb
    opts.seed = int(seed if seed is not None else torch.randint(0, 2 ** 62, (1,)).item()) if perturb > 0 else 0
    opts.white_bkgd = int(bool(getattr(hparams, "white_bkgd", False)))
    opts.precision = L.PRECISIONS[model.precision]
    opts.route = model.route_opts()
    # rendering.py:316-322: sigma noise is drawn per chunk in training mode (torch's generator, like the reference)
    noise_c = noise_f = None
    if model.training and getattr(hparams, "use_sigma_noise", False) and float(getattr(hparams, "sigma_noise_std", 0.0)) > 0:
        std = float(hparams.sigma_noise_std)
        noise_c = torch.randn(N * Sc, dtype=torch.float32, device=dev) * std
        noise_f = torch.randn(N * Sf, dtype=torch.float32, device=dev) * std if Sf > 0 else None
    opts.sigma_noise_coarse, opts.sigma_noise_fine = L.ptr(noise_c), L.ptr(noise_f)
    for flag in ("use_random_background_color", "return_pts", "return_pts_rgb", "return_pts_alpha", "return_sigma", "return_alpha"):
        if getattr(hparams, flag, False):
            raise NotImplementedError(f"hparams.{flag} is not implemented by switch_nerf_b200.rendering.render_rays "
                                      "(debug / visualisation outputs of rendering.py:385-409 outside the hot path)")

    f32 = dict(dtype=torch.float32, device=dev)
    n_chunks_c = -(-N * Sc // chunk) if N > 0 else 0
    n_chunks_f = -(-N * Sf // chunk) if (N > 0 and Sf > 0) else 0
    res = {}
    rgb = torch.empty(N, 3, **f32)
    depth = torch.empty(N, **f32) if (get_depth or get_depth_variance) else None
    var = torch.empty(N, **f32) if get_depth_variance else None
    gl_c = torch.zeros(n_chunks_c, **f32)
    gl_f = torch.zeros(n_chunks_f, **f32) if Sf > 0 else None
    want_gates = bool(getattr(hparams, "moe_return_gates", False))
    mg_c = torch.empty(N, Sc, dtype=torch.int32, device=dev) if want_gates else None
    mg_f = torch.empty(N, Sf, dtype=torch.int32, device=dev) if (want_gates and Sf > 0) else None
    out = L.RenderOut()
    out.rgb, out.depth, out.depth_variance = L.ptr(rgb), L.ptr(depth), L.ptr(var)
    out.gate_loss_coarse, out.gate_loss_fine = L.ptr(gl_c), L.ptr(gl_f)
    out.moe_gates_coarse, out.moe_gates_fine = L.ptr(mg_c), L.ptr(mg_f)
    taps = {}
    if debug_taps:
        taps["raw_coarse"] = torch.empty(N, Sc, 4, **f32)
        out.raw_coarse = L.ptr(taps["raw_coarse"])
        if Sf > 0:
            taps["raw_fine"] = torch.empty(N, Sf, 4, **f32)
            taps["z_fine"] = torch.empty(N, Sf, **f32)
            out.raw_fine, out.z_fine = L.ptr(taps["raw_fine"]), L.ptr(taps["z_fine"])

    lib = L.lib()
    h = model.handle()
    with torch.cuda.device(dev):
        nbytes = lib.snb_render_workspace_bytes(h, N, C.byref(opts))
        ws = L.Workspace.get(nbytes, dev)
        L.check(lib.snb_render_rays(h, L.ptr(rays), L.ptr(idx32), None, N, C.byref(opts), C.byref(out), L.ptr(ws),
                                    ws.numel(), L.stream_handle()))

    # result keys of rendering.py:385-409, 466-494
    res["gate_loss_coarse"] = gl_c
    if want_gates:
        res["moe_gates_coarse"] = mg_c.long().view(N, Sc, 1, 1)
    if Sf > 0:
        res["gate_loss_fine"] = gl_f
        if want_gates:
            res["moe_gates_fine"] = mg_f.long().view(N, Sf, 1, 1)
    res[f"rgb_{typ}"] = rgb
    if get_depth:
        res[f"depth_{typ}"] = depth
    if get_depth_variance:
        res[f"depth_variance_{typ}"] = var
    for k, v in taps.items():
        res[f"_{k}"] = v
    return res, False
