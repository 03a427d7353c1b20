"""Host-side mirror of `switch_nerf.models.nerf.NeRF` (reference models/nerf.py:75-191) in the role the hot path gives
it: the background model behind the unit sphere (`xyz_dim = 4`, get_bg_nerf in models/model_utils.py), evaluated by
`rendering.render_rays` for the rays that leave the foreground sphere.

A parameter container with the reference's module tree -- `xyz_encodings.{i}.0.{weight,bias}`, `xyz_encoding_final`,
`dir_a_encoding.0`, `sigma`, `rgb`, `embedding_a` -- so reference checkpoints (`bg_model_state_dict`) load unchanged;
`forward` hands the chunk to `snb_bg_forward` (fp32 kernels, csrc/snb_bg.cu).  The constructor creates the same
modules in the same order as the reference, so equal seeds give equal initial weights.

Not built (raise at construction): rgb_dim > 3 (spherical harmonics), affine_appearance, pos_dir_dim == 0 -- none of
them is used by the Switch-NeRF configs.
"""
import ctypes as C
from typing import List, Optional

import torch
from torch import nn

from . import _lib as L


class ShiftedSoftplus(nn.Module):
    """models/nerf.py:58-72 marker module: softplus(x - 1); the activation itself runs in the sigma GEMM's epilogue."""

    def __init__(self, beta: int = 1, threshold: int = 20):
        super().__init__()
        assert beta == 1 and threshold == 20, "the kernel implements the reference defaults"


class NeRF(nn.Module):
    def __init__(self, pos_xyz_dim: int, pos_dir_dim: int, layers: int, skip_layers: List[int], layer_dim: int,
                 appearance_dim: int, affine_appearance: bool, appearance_count: int, rgb_dim: int, xyz_dim: int,
                 sigma_activation: nn.Module):
        super().__init__()
        if rgb_dim != 3 or affine_appearance or pos_dir_dim <= 0:
            raise L.SnbError("NeRF (background): rgb_dim == 3, affine_appearance = False, pos_dir_dim > 0 are built")
        if xyz_dim != 4:
            raise L.SnbError("NeRF: only the background role (xyz_dim = 4) is on the hot path; the foreground is NeRFMoE")
        if len(skip_layers) > 1 or any(s <= 0 or s >= layers for s in skip_layers):
            raise L.SnbError("NeRF: one interior skip layer at most")
        if not isinstance(sigma_activation, (nn.ReLU, ShiftedSoftplus)) and type(sigma_activation).__name__ != "ShiftedSoftplus":
            raise L.SnbError("NeRF: sigma_activation must be nn.ReLU or ShiftedSoftplus")
        self.xyz_dim, self.skip_layers = xyz_dim, list(skip_layers)
        self.pos_xyz_dim, self.pos_dir_dim, self.layer_dim = pos_xyz_dim, pos_dir_dim, layer_dim
        self.appearance_dim, self.appearance_count = appearance_dim, appearance_count
        in_xyz = xyz_dim + xyz_dim * pos_xyz_dim * 2
        self.xyz_encodings = nn.ModuleList()
        for i in range(layers):
            fan_in = in_xyz if i == 0 else (layer_dim + in_xyz if i in skip_layers else layer_dim)
            self.xyz_encodings.append(nn.Sequential(nn.Linear(fan_in, layer_dim), nn.ReLU(True)))
        in_dir = 3 + 3 * pos_dir_dim * 2
        self.embedding_a = nn.Embedding(appearance_count, appearance_dim) if appearance_dim > 0 else None
        self.xyz_encoding_final = nn.Linear(layer_dim, layer_dim)
        self.dir_a_encoding = nn.Sequential(nn.Linear(layer_dim + in_dir + appearance_dim, layer_dim // 2), nn.ReLU(True))
        self.sigma = nn.Linear(layer_dim, 1)
        self.sigma_activation = sigma_activation
        self.rgb = nn.Linear(layer_dim // 2, 3)
        self._handle = None
        self._versions = None
        self._device = None

    # ---- device model -------------------------------------------------------------------------
    def _weights(self):
        w = L.BgWeights()
        keep = []

        def p(t):
            t = t.detach().to(torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()
        for i, seq in enumerate(self.xyz_encodings):
            w.w[i], w.b[i] = p(seq[0].weight), p(seq[0].bias)
        w.final_w, w.final_b = p(self.xyz_encoding_final.weight), p(self.xyz_encoding_final.bias)
        w.dir_w, w.dir_b = p(self.dir_a_encoding[0].weight), p(self.dir_a_encoding[0].bias)
        w.sigma_w, w.sigma_b = p(self.sigma.weight), p(self.sigma.bias)
        w.rgb_w, w.rgb_b = p(self.rgb.weight), p(self.rgb.bias)
        w.emb_a = p(self.embedding_a.weight) if self.embedding_a is not None else None
        return w, keep

    def _sync(self):
        dev = self.sigma.weight.device
        if dev.type != "cuda":
            raise L.SnbError("NeRF (background) runs on a CUDA device only (there is no CPU path)")
        versions = tuple(p._version for p in self.parameters()) + tuple(p.data_ptr() for p in self.parameters())
        lib = L.lib()
        if self._handle is not None and self._device != dev:      # the module moved to another GPU: new device copies
            with torch.cuda.device(self._device):
                lib.snb_bg_destroy(self._handle)
            self._handle = None
        with torch.cuda.device(dev):
            if self._handle is None:
                d = L.BgDesc(len(self.xyz_encodings), self.skip_layers[0] if self.skip_layers else -1, self.layer_dim,
                             self.pos_xyz_dim, self.pos_dir_dim, self.appearance_dim, self.appearance_count,
                             0 if isinstance(self.sigma_activation, nn.ReLU) else 1)
                w, keep = self._weights()
                h = C.c_void_p()
                L.check(lib.snb_bg_create(C.byref(d), C.byref(w), L.stream_handle(), C.byref(h)))
                self._handle, self._versions, self._device = h, versions, dev
                torch.cuda.current_stream().synchronize()
            elif versions != self._versions:
                w, keep = self._weights()
                L.check(lib.snb_bg_update(self._handle, C.byref(w), L.stream_handle()))
                self._versions = versions
                torch.cuda.current_stream().synchronize()
        return self._handle

    def mark_dirty(self):
        """Re-upload the weights at the next call (needed after writes through `param.data`, which do not bump
        `Parameter._version`)."""
        self._versions = None

    def __del__(self):
        try:                                     # may run during interpreter shutdown: nothing here is allowed to raise
            h = self.__dict__.get("_handle")
            self.__dict__["_handle"] = None
            if h is not None:
                L.lib().snb_bg_destroy(h)
        except Exception:
            pass

    def forward(self, x: torch.Tensor, sigma_only: bool = False, sigma_noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        if sigma_only:
            raise L.SnbError("NeRF.forward(sigma_only=True) is not on the render path")
        expected = self.xyz_dim + 3 + (1 if self.embedding_a is not None else 0)
        if x.shape[1] != expected:
            raise Exception('Unexpected input shape: {} (expected: {}, xyz_dim: {})'.format(x.shape, expected, self.xyz_dim))
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            _warn_forward_only()
        h = self._sync()
        S = x.shape[0]
        x8 = x.to(torch.float32)
        if x8.shape[1] == 7:
            x8 = torch.cat([x8, x8.new_zeros(S, 1)], 1)
        x8 = x8.contiguous()
        out = torch.empty(S, 4, dtype=torch.float32, device=x.device)
        noise = None if sigma_noise is None else sigma_noise.to(torch.float32).contiguous().view(-1)
        lib = L.lib()
        with torch.cuda.device(x.device):
            nb = lib.snb_bg_workspace_bytes(h, S)
            ws = L.Workspace.get(nb, x.device, tag="bg")
            L.check(lib.snb_bg_forward(h, L.ptr(x8), S, L.ptr(noise) if noise is not None else None, L.ptr(out), L.ptr(ws),
                                       nb, L.stream_handle()))
        return out


_WARNED = False


def _warn_forward_only():
    global _WARNED
    if not _WARNED:
        _WARNED = True
        import warnings
        warnings.warn("switch_nerf_b200.nerf.NeRF: the background model is forward-only; its parameters receive no gradient")
