"""switch_nerf_b200 -- B200-native forward/render hot path of Switch-NeRF behind the reference's operator surface.

Module map (reference -> here):
  switch_nerf.models.nerf_moe                          -> switch_nerf_b200.models.nerf_moe   (NeRFMoE, MipNeRFMoE, get_nerf_moe_inner)
  switch_nerf.modules.tutel_moe_ext.tutel_moe_nobatch  -> switch_nerf_b200.modules.tutel_moe_ext.tutel_moe_nobatch (moe_layer, SingleExpert, fast_cumsum_sub_one)
  switch_nerf.rendering / switch_nerf.rendering_mip    -> switch_nerf_b200.rendering / switch_nerf_b200.rendering_mip (render_rays)
"""
import importlib
import sys

_ALIASES = {
    "switch_nerf.models.nerf_moe": "switch_nerf_b200.models.nerf_moe",
    "switch_nerf.modules.tutel_moe_ext.tutel_moe_nobatch": "switch_nerf_b200.modules.tutel_moe_ext.tutel_moe_nobatch",
    "switch_nerf.rendering": "switch_nerf_b200.rendering",
    "switch_nerf.rendering_mip": "switch_nerf_b200.rendering_mip",
}


def install_as_switch_nerf(names=None):
    """Register the drop-in modules under the reference's import paths, so an unmodified caller
    (`from switch_nerf.rendering import render_rays`, `from switch_nerf.models.nerf_moe import get_nerf_moe_inner`)
    gets the fused path.  `names`: subset of the reference module paths (default: all four).  Returns the mapping."""
    done = {}
    for ref_name, mine in _ALIASES.items():
        if names is not None and ref_name not in names:
            continue
        sys.modules[ref_name] = done[ref_name] = importlib.import_module(mine)
    return done
