"""Host-side mirror of `switch_nerf.ray_utils` (reference ray_utils.py:6-84) over the C ABI: camera rays of one image
are generated on the device by one kernel (`snb_get_rays`), not by the meshgrid / matmul / boolean-mask chain of torch
ops -- SURVEY.md 8f-2."""
import ctypes as C
from typing import List, Optional

import torch

from . import _lib as L


def get_rays_for_image(W: int, H: int, fx: float, fy: float, cx: float, cy: float, center_pixels: bool, c2w: torch.Tensor,
                       near: float, far: float, ray_altitude_range: Optional[List[float]]) -> torch.Tensor:
    """== get_rays(get_ray_directions(W, H, fx, fy, cx, cy, center_pixels, device), c2w, near, far, ray_altitude_range):
    rays [H, W, 8] = [origin, unit direction, near, far] (ray_utils.py:6-65)."""
    if not c2w.is_cuda:
        raise L.SnbError("get_rays_for_image: c2w must be a CUDA tensor (switch_nerf_b200 has no CPU path)")
    c = c2w.detach().to(torch.float32).contiguous()
    if c.shape != (3, 4):
        raise L.SnbError(f"get_rays_for_image: c2w must be [3,4], got {tuple(c.shape)}")
    rays = torch.empty(H, W, 8, dtype=torch.float32, device=c.device)
    alt = None
    if ray_altitude_range is not None:
        alt = (C.c_float * 2)(float(ray_altitude_range[0]), float(ray_altitude_range[1]))
    with torch.cuda.device(c.device):
        L.check(L.lib().snb_get_rays(int(W), int(H), float(fx), float(fy), float(cx), float(cy), int(bool(center_pixels)),
                                     L.ptr(c), float(near), float(far), alt, L.ptr(rays), L.stream_handle()))
    return rays
