"""Host-side mirror of the evaluation tiling of the reference (`Runner.render_image`, runner.py:2835-2885, and
`Runner.render_image_blocknerf`, :2887-2940): the rays of one image are generated on the device
(`snb_get_rays`) and rendered in batches of `hparams.image_pixel_batch_size`; the per-batch `results` are concatenated
key by key exactly as the reference does (per-ray tensors along the ray axis, the per-chunk gate losses along the chunk
axis).  The only difference: batches stay on the device until the end (one D2H copy per key instead of one per batch).
"""
from argparse import Namespace
from typing import Dict, List, Optional, Tuple

import torch

from .ray_utils import get_rays_for_image
from .rendering import render_rays
from .rendering_mip import render_rays as render_rays_mip


def _concat(results: Dict[str, List[torch.Tensor]], to_cpu: bool) -> Dict[str, torch.Tensor]:
    out = {}
    for key, vals in results.items():
        v = torch.cat([t if t.dim() > 0 else t.view(1) for t in vals])
        out[key] = v.cpu() if to_cpu else v
    return out


def render_image(nerf, W: int, H: int, intrinsics, c2w: torch.Tensor, image_index: int, hparams: Namespace, near: float,
                 far: float, ray_altitude_range: Optional[List[float]] = None, bg_nerf=None, sphere_center=None,
                 sphere_radius=None, to_cpu: bool = True) -> Tuple[Dict[str, torch.Tensor], torch.Tensor]:
    """== Runner.render_image(metadata) with metadata = (W, H, intrinsics = [fx, fy, cx, cy], c2w, image_index)."""
    rays = get_rays_for_image(W, H, float(intrinsics[0]), float(intrinsics[1]), float(intrinsics[2]), float(intrinsics[3]),
                              bool(getattr(hparams, "center_pixels", True)), c2w, near, far, ray_altitude_range).view(-1, 8)
    image_indices = None
    if int(getattr(hparams, "appearance_dim", 0)) > 0:
        image_indices = torch.full((rays.shape[0],), int(image_index), dtype=torch.int32, device=rays.device)
    results: Dict[str, List[torch.Tensor]] = {}
    bs = int(hparams.image_pixel_batch_size)
    with torch.no_grad():
        for i in range(0, rays.shape[0], bs):
            batch, _ = render_rays(nerf, bg_nerf, rays[i:i + bs], None if image_indices is None else image_indices[i:i + bs],
                                   hparams, sphere_center, sphere_radius, True, False, True)
            for key, value in batch.items():
                results.setdefault(key, []).append(value)
    return _concat(results, to_cpu), rays


def render_image_blocknerf(nerf, rays: torch.Tensor, radii: Optional[torch.Tensor], image_indices: Optional[torch.Tensor],
                           hparams: Namespace, to_cpu: bool = True) -> Tuple[Dict[str, torch.Tensor], torch.Tensor]:
    """== Runner.render_image_blocknerf(rays, radii, image_indices): pre-computed rays (Block-NeRF / Mission Bay data), the
    mip renderer when radii are given."""
    dev = next(nerf.parameters()).device
    rays = rays.reshape(-1, 8).to(dev, non_blocking=True)
    radii = None if radii is None else radii.reshape(-1, 1).to(dev, non_blocking=True)
    image_indices = None if image_indices is None else image_indices.reshape(-1).to(dev, non_blocking=True)
    use_idx = int(getattr(hparams, "appearance_dim", 0)) > 0 and image_indices is not None
    results: Dict[str, List[torch.Tensor]] = {}
    bs = int(hparams.image_pixel_batch_size)
    with torch.no_grad():
        for i in range(0, rays.shape[0], bs):
            idx = image_indices[i:i + bs] if use_idx else None
            if radii is None:
                batch, _ = render_rays(nerf, None, rays[i:i + bs], idx, hparams, None, None, True, False, False)
            else:
                batch, _ = render_rays_mip(nerf, rays[i:i + bs], radii[i:i + bs], idx, hparams, True, False)
            for key, value in batch.items():
                results.setdefault(key, []).append(value)
    return _concat(results, to_cpu), rays
