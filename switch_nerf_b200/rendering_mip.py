"""Host-side mirror of `switch_nerf.rendering_mip` (reference rendering_mip.py:133-174) over the C ABI:
`render_rays(nerf, rays, radii, image_indices, hparams, get_depth, get_depth_variance)` for MipNeRFMoE
models (Mission Bay / Bungee configs).  Eval-mode sampling (hparams.perturb = 0) only; the model chunks
run on the fp32 CUDA path (the tcgen05 path is specialised for the width-256 NeRFMoE topology)."""
import ctypes as C
from argparse import Namespace
from typing import Dict, Optional, Tuple

import torch

from . import _lib as L
from .nerf_moe import MipNeRFMoE


def render_rays(nerf, rays: torch.Tensor, radii: torch.Tensor, image_indices: Optional[torch.Tensor],
                hparams: Namespace, get_depth: bool = True, get_depth_variance: bool = True,
                debug_taps: bool = False, seed: Optional[int] = None,
                deterministic_eval: bool = False) -> Tuple[Dict[str, torch.Tensor], bool]:
    """reference rendering_mip.render_rays (133-174).  Like the reference, the coarse depths are perturbed only in
    training mode (:147), while the fine resampling is randomised whenever `hparams.perturb` is non-zero -- in eval too
    (:225, `randomized=hparams.perturb`; the reference default is perturb = 1.0).  `deterministic_eval=True` turns that
    quirk off for reproducible evaluation; `seed` fixes the stratified jitter (default: drawn from torch's generator)."""
    model = nerf.module if hasattr(nerf, "module") and isinstance(nerf.module, MipNeRFMoE) else nerf
    if not isinstance(model, MipNeRFMoE):
        raise L.SnbError("rendering_mip.render_rays needs a switch_nerf_b200.nerf_moe.MipNeRFMoE model")
    rays = L.require_cuda_f32(rays, "rays", cols=8)
    N, dev = rays.shape[0], rays.device
    radii = L.require_cuda_f32(radii.reshape(-1), "radii")
    Sc, Sf = int(hparams.coarse_samples), int(hparams.fine_samples)
    chunk = int(hparams.model_chunk_size)
    idx32 = None if image_indices is None else image_indices.to(device=dev, dtype=torch.int32).contiguous()
    opts = L.RenderOpts()
    opts.coarse_samples, opts.fine_samples, opts.model_chunk_size = Sc, Sf, chunk
    hp_perturb = float(getattr(hparams, "perturb", 0.0))
    opts.perturb = hp_perturb if model.training else 0.0
    opts.resample_randomized = int(hp_perturb != 0.0 and not (deterministic_eval and not model.training))
    need_seed = opts.perturb > 0 or opts.resample_randomized
    opts.seed = int(seed if seed is not None else torch.randint(0, 2 ** 62, (1,)).item()) if need_seed else 0
    opts.white_bkgd = int(bool(getattr(hparams, "white_bkgd", False)))
    opts.precision = L.PRECISIONS[model.precision]
    opts.route = model.route_opts()
    f32 = dict(dtype=torch.float32, device=dev)
    typ = "fine" if Sf > 0 else "coarse"
    rgb, rgb_c = torch.empty(N, 3, **f32), torch.empty(N, 3, **f32)
    depth = torch.empty(N, **f32) if (get_depth or get_depth_variance) else None
    var = torch.empty(N, **f32) if get_depth_variance else None
    nc_c = -(-N * (Sc - 1) // chunk) if N > 0 else 0
    nc_f = -(-N * (Sf - 1) // chunk) if (N > 0 and Sf > 0) else 0
    gl_c, gl_f = torch.zeros(nc_c, **f32), (torch.zeros(nc_f, **f32) if Sf > 0 else None)
    want_gates = bool(getattr(hparams, "moe_return_gates", False))
    mg_c = torch.empty(N, Sc - 1, dtype=torch.int32, device=dev) if want_gates else None
    mg_f = torch.empty(N, Sf - 1, dtype=torch.int32, device=dev) if (want_gates and Sf > 0) else None
    out = L.RenderOut()
    out.rgb, out.rgb_coarse, out.depth, out.depth_variance = L.ptr(rgb), L.ptr(rgb_c), L.ptr(depth), L.ptr(var)
    out.gate_loss_coarse, out.gate_loss_fine = L.ptr(gl_c), L.ptr(gl_f)
    out.moe_gates_coarse, out.moe_gates_fine = L.ptr(mg_c), L.ptr(mg_f)
    taps = {}
    if debug_taps and Sf > 0:
        taps["z_fine"] = torch.empty(N, Sf, **f32)
        out.z_fine = L.ptr(taps["z_fine"])
    pad = getattr(hparams, "rgb_padding", None)
    lib, h = L.lib(), model.handle()
    with torch.cuda.device(dev):
        nbytes = lib.snb_render_mip_workspace_bytes(h, N, C.byref(opts))
        ws = L.Workspace.get(nbytes, dev)
        L.check(lib.snb_render_rays_mip(h, L.ptr(rays), L.ptr(radii), L.ptr(idx32), None, N, C.byref(opts),
                                        float(getattr(hparams, "weights_resample_padding", 0.01)),
                                        -1.0 if pad is None else float(pad), C.byref(out), L.ptr(ws), ws.numel(),
                                        L.stream_handle()))
    res = {"gate_loss_coarse": gl_c}
    if want_gates:
        res["moe_gates_coarse"] = mg_c.long().view(N, Sc - 1, 1, 1)
    if Sf > 0:
        res["rgb_coarse"] = rgb_c
        res["gate_loss_fine"] = gl_f
        if want_gates:
            res["moe_gates_fine"] = mg_f.long().view(N, Sf - 1, 1, 1)
    res[f"rgb_{typ}"] = rgb
    if get_depth:
        res[f"depth_{typ}"] = depth
    if get_depth_variance:
        res[f"depth_variance_{typ}"] = var
    for k, v in taps.items():
        res[f"_{k}"] = v
    return res, False
