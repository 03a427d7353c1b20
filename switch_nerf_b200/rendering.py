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
    typ = "fine" if Sf > 0 else "coarse"
    want_grad = torch.is_grad_enabled() and any(p.requires_grad for p in model.parameters())

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
    if debug_taps or want_grad:
        taps["z_coarse"] = torch.empty(N, Sc, **f32)
        out.z_coarse = L.ptr(taps["z_coarse"])
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

    if want_grad:
        # attach the autograd graph: rgb and the per-chunk gate losses are functions of the parameters (depth and its
        # variance are computed under no_grad in the reference too, rendering.py:479-494)
        if getattr(hparams, "white_bkgd", False):
            raise NotImplementedError("backward with white_bkgd is not implemented")
        if model.layers["0"].moe_no_batch:
            raise NotImplementedError("backward implements the capacity (batched) dispatch, the reference's training mode")
        plist = [p for _, _, p in model._grad_params()]
        saved = dict(model=model, rays=rays, idx32=idx32, N=N, Sc=Sc, Sf=Sf, chunk=chunk, taps=taps, noise_c=noise_c, noise_f=noise_f)
        rgb, gl_c, gl_f_t = _RenderGrad.apply(saved, rgb, gl_c, gl_f if gl_f is not None else gl_c.new_zeros(0), *plist)
        if gl_f is not None:
            gl_f = gl_f_t
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
    if debug_taps:
        for k, v in taps.items():
            res[f"_{k}"] = v
    return res, False


class _RenderGrad(torch.autograd.Function):
    """Backward of render_rays for the parameters (what runner.py:677-690 calls through loss.backward()):
    composite^T (snb_composite_backward; the fine pass through the sorted merge of rendering.py:419-431), then the
    model-chunk backward of every chunk of both passes (snb_moe_backward: heads, combine^T + gate-value gradient, expert
    dgrad / wgrad with the skip, dispatch^T, l_aux term, softmax / LayerNorm / gate MLP / xyz layer, embedding
    scatter).  fp32 CUDA kernels that recompute the forward intermediates of a chunk; with a bf16 forward this is an
    fp32 backward of the same function.  The fine sample positions come from detached coarse weights (rendering.py:240),
    so nothing flows through them."""

    @staticmethod
    def forward(ctx, saved, rgb, gl_c, gl_f, *params):
        ctx.saved = saved
        ctx.n_params = len(params)
        return rgb.view_as(rgb), gl_c.view_as(gl_c), gl_f.view_as(gl_f)

    @staticmethod
    def backward(ctx, d_rgb, d_gl_c, d_gl_f):
        sv = ctx.saved
        model, rays, idx32, N, Sc, Sf, chunk, taps = (sv[k] for k in ("model", "rays", "idx32", "N", "Sc", "Sf", "chunk", "taps"))
        dev = rays.device
        lib = L.lib()
        f32 = dict(dtype=torch.float32, device=dev)
        d_rgb = torch.zeros(N, 3, **f32) if d_rgb is None else d_rgb.contiguous().float()
        zc, raw_c = taps["z_coarse"], taps["raw_coarse"]
        with torch.no_grad(), torch.cuda.device(dev):
            def comp_bwd(z, raw):
                d_raw = torch.empty_like(raw)
                L.check(lib.snb_composite_backward(L.ptr(z), L.ptr(raw), None, z.shape[0], z.shape[1], L.ptr(d_rgb),
                                                   L.ptr(d_raw), L.stream_handle()))
                return d_raw
            if Sf > 0:
                zf, raw_f = taps["z_fine"], taps["raw_fine"]
                z_all, order = torch.sort(torch.cat([zf, zc], -1), dim=-1, stable=True)               # rendering.py:421
                raw_all = torch.gather(torch.cat([raw_f, raw_c], 1), 1, order.unsqueeze(-1).expand(-1, -1, 4)).contiguous()
                d_all = comp_bwd(z_all.contiguous(), raw_all)
                d_cat = torch.zeros_like(d_all).scatter_(1, order.unsqueeze(-1).expand(-1, -1, 4), d_all)
                passes = [(zf, d_cat[:, :Sf].contiguous(), d_gl_f, sv["noise_f"]), (zc, d_cat[:, Sf:].contiguous(), d_gl_c, sv["noise_c"])]
            else:
                passes = [(zc, comp_bwd(zc, raw_c), d_gl_c, sv["noise_c"])]
            fields = model._grad_params()
            grads = [torch.zeros_like(p, dtype=torch.float32) if p.requires_grad else None for _, _, p in fields]
            G = L.Weights()
            for (name, i, _), g in zip(fields, grads):
                if g is None:
                    continue
                if i is None:
                    setattr(G, name, g.data_ptr())
                else:
                    getattr(G, name)[i] = g.data_ptr()
            h = model.handle()
            opts = model.route_opts()
            nbytes = lib.snb_moe_backward_workspace_bytes(h, min(chunk, N * max(Sc, Sf)), opts.capacity_factor)
            ws = L.Workspace.get(nbytes, dev, tag="backward")
            o, d = rays[:, None, 0:3], rays[:, None, 3:6]
            img = (idx32 if idx32 is not None else torch.zeros(N, dtype=torch.int32, device=dev)).float()
            for z, d_raw, d_gl, noise in passes:
                Sn = z.shape[1]
                x = torch.cat([o + d * z.unsqueeze(-1), d.expand(N, Sn, 3), img.view(N, 1, 1).expand(N, Sn, 1)], -1).reshape(N * Sn, 7).contiguous()
                d_out = d_raw.reshape(N * Sn, 4)
                d_gl = None if d_gl is None else d_gl.contiguous().float()
                for ci, i0 in enumerate(range(0, N * Sn, chunk)):
                    rows = min(chunk, N * Sn - i0)
                    L.check(lib.snb_moe_backward(h, L.ptr(x[i0:i0 + rows]), rows, L.ptr(noise[i0:i0 + rows]) if noise is not None else None,
                                                 C.byref(opts), L.ptr(d_out[i0:i0 + rows]),
                                                 L.ptr(d_gl[ci:ci + 1]) if d_gl is not None else None, C.byref(G), L.ptr(ws),
                                                 ws.numel(), L.stream_handle()))
        out = [None if g is None else g.to(p.dtype) for g, (_, _, p) in zip(grads, fields)]
        return (None, None, None, None, *out)
