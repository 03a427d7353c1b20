"""`switch_nerf.models.nerf_moe` import surface (reference models/nerf_moe.py:16-49, 103-455, 458-810, 1004-1041):
the same names at the same module path, so `sys.modules["switch_nerf.models.nerf_moe"] = switch_nerf_b200.models.nerf_moe`
(or `switch_nerf_b200.install_as_switch_nerf()`) swaps the fused path in under an unmodified caller."""
from ..nerf_moe import Mlp, MipNeRFMoE, NeRFMoE, get_nerf_moe_inner  # noqa: F401
