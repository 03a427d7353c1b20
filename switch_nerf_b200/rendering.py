"""Host-side mirror of `switch_nerf.rendering` (reference rendering.py:15-196) over the C ABI.

`render_rays` keeps the reference signature and `results` keys so `Runner._training_step` /
`Runner.render_image` (runner.py:1077-1123, 2835-2885) call it unchanged.  The whole ray pipeline
(coarse depths, point generation, chunked model evaluation, transmittance scan, inverse-CDF
resampling, fine pass, sorted merge and composite) is one `snb_render_rays` call; torch only
owns the output tensors.

Scope: foreground NeRFMoE model, `use_cascade=False`, `pos_dir_dim > 0`,
`sh_deg=None` -- the path of the Building-family configs.  With `bg_nerf` (rendering.py:33-80,
104-146) the rays that leave the foreground sphere are continued through the background model: sphere
intersection, inverted-sphere points, the model chunks, resampling and the composites are library calls
(`snb_intersect_sphere`, `snb_depth2pts_outside`, `snb_bg_forward`, `snb_sample_pdf`, `snb_composite`); torch
does the index glue between them (ray subset, flips, the descending merge).  The bg branch is forward-only.
"""
import ctypes as C
from argparse import Namespace
from typing import Dict, Optional, Tuple

import torch

from . import _lib as L
from .nerf_moe import NeRFMoE
from .nerf import NeRF


def _unwrap(nerf):
    return nerf.module if hasattr(nerf, "module") and isinstance(nerf.module, NeRFMoE) else nerf


def render_rays(nerf, bg_nerf, rays: torch.Tensor, image_indices: Optional[torch.Tensor], hparams: Namespace,
                sphere_center: Optional[torch.Tensor] = None, sphere_radius: Optional[torch.Tensor] = None,
                get_depth: bool = True, get_depth_variance: bool = True, get_bg_fg_rgb: bool = False,
                debug_taps: bool = False, seed: Optional[int] = None) -> Tuple[Dict[str, torch.Tensor], bool]:
    model = _unwrap(nerf)
    if not isinstance(model, NeRFMoE):
        raise L.SnbError("render_rays needs a switch_nerf_b200.nerf_moe.NeRFMoE model")
    bg_model = None
    if bg_nerf is not None:
        bg_model = bg_nerf.module if hasattr(bg_nerf, "module") and isinstance(bg_nerf.module, NeRF) else bg_nerf
        if not isinstance(bg_model, NeRF):
            raise L.SnbError("render_rays: bg_nerf must be a switch_nerf_b200.nerf.NeRF (bg_use_moe is not built)")
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
    lib = L.lib()

    # ---- background model for the rays that leave the sphere (rendering.py:33-80) -------------------------------
    last_delta = with_bg = bg_res = None
    if bg_model is not None:
        if want_grad:
            raise NotImplementedError("render_rays with bg_nerf is forward-only: call it under torch.no_grad()")
        c = None if sphere_center is None else sphere_center.to(device=dev, dtype=torch.float32).contiguous()
        rad = None if sphere_radius is None else sphere_radius.to(device=dev, dtype=torch.float32).contiguous()
        fg_far = torch.empty(N, dtype=torch.float32, device=dev)
        bad = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            L.check(lib.snb_intersect_sphere(L.ptr(rays), N, L.ptr(c), L.ptr(rad), L.ptr(fg_far), L.ptr(bad), L.stream_handle()))
        if int(bad.item()):
            raise Exception('Not all your cameras are bounded by the unit sphere; please make sure the cameras are '
                            'normalized properly!')
        fg_far = torch.maximum(fg_far, rays[:, 6])
        with_bg = torch.nonzero(rays[:, 7] > fg_far).view(-1)
        if with_bg.numel() > 0:
            last_delta = torch.full((N,), 1e10, dtype=torch.float32, device=dev)
            last_delta[with_bg] = fg_far[with_bg]
            rays_bg = rays[with_bg].contiguous()
            rays = rays.clone()
            rays[:, 7] = torch.minimum(rays[:, 7], fg_far)
            bg_res = _bg_results(bg_model, rays_bg, None if image_indices is None else image_indices.to(dev)[with_bg], hparams,
                                 c, rad, perturb, get_depth, get_depth_variance)

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

    lam = None
    if bg_model is not None:
        lam = torch.empty(N, **f32)
        out.bg_lambda = L.ptr(lam)
        opts.last_delta_minus_zmax = 1
    h = model.handle()
    with torch.cuda.device(dev):
        nbytes = lib.snb_render_workspace_bytes(h, N, C.byref(opts))
        ws = L.Workspace.get(nbytes, dev)
        L.check(lib.snb_render_rays(h, L.ptr(rays), L.ptr(idx32), L.ptr(last_delta), N, C.byref(opts), C.byref(out), L.ptr(ws),
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
    if bg_model is not None:
        res[f"bg_lambda_{typ}"] = lam                                   # rendering.py:456-457
        # rendering.py:104-146: fg + bg_lambda * bg for the composited keys
        for key in ("rgb", "depth"):
            name = f"{key}_{typ}"
            if name not in res:
                continue
            val = res[name]
            bg_val = torch.zeros_like(val)
            if bg_res is not None:
                mult = lam[with_bg]
                bg_val[with_bg] = bg_res[name] * (mult.unsqueeze(-1) if val.dim() > 1 else mult)
            if get_bg_fg_rgb:
                res[f"fg_{name}"], res[f"bg_{name}"] = val, bg_val
            if bg_res is not None:
                res[name] = val + bg_val
        if debug_taps and bg_res is not None:
            res["_rays_with_bg"] = with_bg
            for k, v in bg_res.items():
                res[f"_bg_{k}"] = v
    return res, bg_res is not None


def _bg_chunks(bg_model, pts, rays_bg, idx, hparams):
    """rendering.py:300-383 for the background model: rows [pts(4), dir(3), image index] in chunks of model_chunk_size."""
    Nb, S, _ = pts.shape
    d = rays_bg[:, None, 3:6].expand(Nb, S, 3)
    cols = [pts, d, (idx.float() if idx is not None else pts.new_zeros(Nb)).view(Nb, 1, 1).expand(Nb, S, 1)]
    x = torch.cat(cols, -1).reshape(Nb * S, 8)
    if bg_model.embedding_a is None:
        x = x[:, :7]
    chunk = int(hparams.model_chunk_size)
    outs = []
    for i in range(0, Nb * S, chunk):
        xc = x[i:i + chunk]
        noise = None
        if bg_model.training and getattr(hparams, "use_sigma_noise", False) and float(getattr(hparams, "sigma_noise_std", 0.0)) > 0:
            noise = torch.randn(len(xc), 1, device=xc.device) * float(hparams.sigma_noise_std)
        outs.append(bg_model(xc, sigma_noise=noise))
    return torch.cat(outs, 0).view(Nb, S, 4)


def _bg_results(bg_model, rays_bg, idx, hparams, c, rad, perturb, get_depth, get_depth_variance):
    """`_get_results(nerf=bg_nerf, flip=True, last_delta=1e10, depth_real=...)` of rendering.py:55-77.  The background
    samples are inverse distances in [0, 1]; the reference composites them in descending order (`flip`) with deltas
    z[i] - z[i+1], which is the ascending composite of -z: that is how `snb_composite` is called here.  Reproduced as
    the reference computes it, including that `depth_real` of the coarse level is NOT flipped with its samples
    (rendering.py:291-294 flips xyz and z_vals only) and that the resampling pdf pairs the flipped weights with the
    unflipped bins (rendering.py:238-241)."""
    lib = L.lib()
    dev = rays_bg.device
    Nb = rays_bg.shape[0]
    Sb, Sf = int(hparams.coarse_samples) // 2, int(hparams.fine_samples)
    f32 = dict(dtype=torch.float32, device=dev)
    z = torch.linspace(0, 1, Sb, device=dev).expand(Nb, Sb)
    if perturb > 0:                                                     # rendering.py:573-584
        mid = 0.5 * (z[:, :-1] + z[:, 1:])
        upper, lower = torch.cat([mid, z[:, -1:]], -1), torch.cat([z[:, :1], mid], -1)
        z = lower + (upper - lower) * (perturb * torch.rand_like(z))
    z = z.contiguous()

    def sphere_pts(zv):
        S = zv.shape[1]
        pts, real = torch.empty(Nb, S, 4, **f32), torch.empty(Nb, S, **f32)
        with torch.cuda.device(dev):
            L.check(lib.snb_depth2pts_outside(L.ptr(rays_bg), L.ptr(c), L.ptr(rad), L.ptr(zv), Nb, S, L.ptr(pts), L.ptr(real),
                                              L.stream_handle()))
        return pts, real

    def composite(z_desc, raw, want_rgb):
        S = z_desc.shape[1]
        w = torch.empty(Nb, S, **f32)
        rgb = torch.empty(Nb, 3, **f32) if want_rgb else None
        zneg = (-z_desc).contiguous()
        with torch.cuda.device(dev):
            L.check(lib.snb_composite(L.ptr(zneg), L.ptr(raw), None, Nb, S, int(bool(getattr(hparams, "white_bkgd", False))) if want_rgb else 0,
                                      L.ptr(rgb), None, None, None, L.ptr(w), L.stream_handle()))
        return w, rgb

    def depth_keys(res, typ, w, real, z_desc):
        if get_depth or get_depth_variance:
            dm = (w * real).sum(1)
            if get_depth:
                res[f"depth_{typ}"] = dm
            if get_depth_variance:
                res[f"depth_variance_{typ}"] = (w * (z_desc - dm.unsqueeze(1)).square()).sum(-1)

    res = {}
    pts, real_c = sphere_pts(z)
    z_c = torch.flip(z, dims=[-1]).contiguous()
    raw_c = _bg_chunks(bg_model, torch.flip(pts, dims=[-2]), rays_bg, idx, hparams).contiguous()
    w_c, rgb_c = composite(z_c, raw_c, Sf == 0)
    if Sf == 0:
        res["rgb_coarse"] = rgb_c
        depth_keys(res, "coarse", w_c, real_c, z_c)
        return res
    nf = Sf // 2
    zmid = (0.5 * (z[:, :-1] + z[:, 1:])).contiguous()
    wmid = w_c[:, 1:-1].contiguous()
    u = None if perturb == 0 else torch.rand(Nb, nf, **f32)
    z_f = torch.empty(Nb, nf, **f32)
    with torch.cuda.device(dev):
        L.check(lib.snb_sample_pdf(L.ptr(zmid), L.ptr(wmid), L.ptr(u), Nb, Sb - 2, nf, L.ptr(z_f), L.stream_handle()))
    pts_f, real_f = sphere_pts(z_f)
    raw_f = _bg_chunks(bg_model, pts_f, rays_bg, idx, hparams)
    z_all, order = torch.sort(torch.cat([z_f, z_c], -1), -1, descending=True)                   # rendering.py:421
    raw_all = torch.gather(torch.cat([raw_f, raw_c], 1), 1, order.unsqueeze(-1).expand(-1, -1, 4)).contiguous()
    real_all = torch.gather(torch.cat([real_f, real_c], 1), 1, order)
    w, rgb = composite(z_all, raw_all, True)
    res["rgb_fine"] = rgb
    depth_keys(res, "fine", w, real_all, z_all)
    return res


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
