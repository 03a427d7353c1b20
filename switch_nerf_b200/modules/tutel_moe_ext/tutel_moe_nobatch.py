"""`switch_nerf.modules.tutel_moe_ext.tutel_moe_nobatch` import surface (reference tutel_moe_nobatch.py:1-10 re-exports
`moe_layer`, `SingleExpert`, `fast_cumsum_sub_one` and the dispatcher)."""
from ...nerf_moe import MOELayer, SingleExpert, TopKGate, fast_cumsum_sub_one, moe_layer  # noqa: F401
