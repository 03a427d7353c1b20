"""Import surface of `switch_nerf.modules.tutel_moe_ext` (the Tutel-backed MoE layer of the reference)."""
